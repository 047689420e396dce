// NRCudaAdapter.cpp — the CUDA backend as an NRenderer render component.
//
// Drop-in for the reference's CPU components behind the unchanged plugin API:
//   class Adapter : public RenderComponent { void render(SharedScene) }   (code/include/component/RenderComponent.hpp:12-18)
//   REGISTER_RENDERER(Name, description, Class)                           (code/include/component/Component.hpp:23-34)
// The reference registers one component per shared library (fixed-name static in the macro), so this
// file is compiled three times:  -DNRCU_PLUGIN_MODE=0 -> "CudaRayCast"           (replaces RayCast,          ray_cast/src/Adapter.cpp:11-34)
//                                -DNRCU_PLUGIN_MODE=1 -> "CudaSimplePathTracer"  (replaces SimplePathTracer, simple_path_tracing/src/Adapter.cpp:13-30)
//                                -DNRCU_PLUGIN_MODE=2 -> "CudaAccPathTracer"     (replaces AccPathTracer,    acc_path_tracing/src/Adapter.cpp:13-30)
//                                -DNRCU_PLUGIN_MODE=3 -> "CudaMetropolisLightTransport" (counterpart of MetropolisLightTransport,
//                                                        metropolis_light_transport/src/Adapter.cpp:13-36: nrcu_render_mlt, the reference MLT's tone map)
// render() never throws and treats the Scene as read-only (the reference components mutate it in
// place); the frame is published through getServer().screen.set exactly like ray_cast/src/Adapter.cpp:15-19.
// It is compiled by g++ against the reference's headers and talks to the kernels only through the
// C ABI in include/nrcu.h.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "server/Server.hpp"
#include "component/RenderComponent.hpp"

#include "nrcu.h"
#include "scene_bridge.hpp"

#ifndef NRCU_PLUGIN_MODE
#define NRCU_PLUGIN_MODE 2
#endif
// the scene mode the component uploads with: the Metropolis component samples paths like SimplePathTracer's closest hit
// (brute-force semantics, no leaf gate: metropolis_light_transport/src/Metropolis.cpp:143-177) with AccPathTracer's materials
#if NRCU_PLUGIN_MODE == 3
#define NRCU_SCENE_MODE NRCU_MODE_ACC
#else
#define NRCU_SCENE_MODE NRCU_PLUGIN_MODE
#endif

using namespace NRenderer;

namespace NRCuda
{
    // One CUDA context per process, created on first use and reused by every render
    // (the GUI runs render() on a fresh detached thread per click, ComponentManager.hpp:41-64).
    struct SharedContext {
        std::mutex mtx;
        nrcu_ctx* ctx = nullptr;               // primary device: NRCU_DEVICE (default 0)
        int ctx_device = -1;
        std::map<int, nrcu_ctx*> extra;        // further devices when NRCU_DEVICES > 1, keyed by DEVICE INDEX (sample slices, nrcu_render_multi)
        // The frame buffer the component owns (ray_cast/src/RayCastRenderer.cpp:7-10 allocates one per render with new[]):
        // page-locked and kept across renders, so that the frame comes back in one DMA; plain memory if pinning fails.
        RGBA* frame = nullptr; size_t frame_pixels = 0; bool frame_pinned = false;
        RGBA* frameBuffer(size_t n) {
            if (frame && frame_pixels >= n) return frame;
            releaseFrame();
            frame = static_cast<RGBA*>(nrcu_host_alloc(n * sizeof(RGBA)));
            frame_pinned = frame != nullptr;
            if (!frame) frame = new RGBA[n];
            frame_pixels = n;
            return frame;
        }
        void releaseFrame() { if (frame) { if (frame_pinned) nrcu_host_free(frame); else delete[] frame; } frame = nullptr; frame_pixels = 0; }
        ~SharedContext() { releaseFrame(); for (auto& kv : extra) nrcu_destroy(kv.second); if (ctx) nrcu_destroy(ctx); }
    };
    static SharedContext& shared() { static SharedContext s; return s; }

    static void publishBlack(unsigned w, unsigned h) {
        std::vector<RGBA> px((size_t)w * h, RGBA{0, 0, 0, 1});
        getServer().screen.set(px.data(), (int)w, (int)h);
    }

    class Adapter : public RenderComponent
    {
        void render(SharedScene spScene) override {
            auto& logger = getServer().logger;
            if (!spScene) { logger.error("NRCuda: no scene (SceneBuilder::build failed?)"); return; }
            const unsigned w = spScene->renderOption.width, h = spScene->renderOption.height;
            try {
                auto t0 = std::chrono::steady_clock::now();
                std::string warn;
                nrb200::FlatScene flat = nrb200::flatten(*spScene, &warn);
                if (!warn.empty()) logger.warning("NRCuda: " + warn);
                nrcu_scene view = flat.view();

                auto& sh = shared();
                std::lock_guard<std::mutex> lock(sh.mtx);
                int dev = 0;
                if (const char* e = std::getenv("NRCU_DEVICE")) dev = std::atoi(e);
                if (sh.ctx && sh.ctx_device != dev) {   // NRCU_DEVICE changed between two renders: rebind, do not keep a stale device
                    nrcu_destroy(sh.ctx); sh.ctx = nullptr;
                    auto it = sh.extra.find(dev);
                    if (it != sh.extra.end()) { nrcu_destroy(it->second); sh.extra.erase(it); }
                }
                if (!sh.ctx) {
                    if (nrcu_create(dev, &sh.ctx) != NRCU_OK) {
                        logger.error(std::string("NRCuda: ") + nrcu_last_error(nullptr));
                        publishBlack(w, h);
                        return;
                    }
                    sh.ctx_device = dev;
                }
                if (nrcu_upload_scene(sh.ctx, &view, NRCU_SCENE_MODE) != NRCU_OK) {
                    logger.error(std::string("NRCuda: ") + nrcu_last_error(sh.ctx));
                    publishBlack(w, h);
                    return;
                }
                // NRCU_DEVICES=N (or "all"): split the samples of the frame over N GPUs of this box.  A device that
                // cannot be opened or refuses the scene is skipped (and said so); the next device index is tried instead.
                std::vector<nrcu_ctx*> devs{sh.ctx};
                std::vector<int> dev_ids{dev};
                if (const char* e = std::getenv("NRCU_DEVICES")) {
                    const int n_dev = nrcu_device_count();
                    const int want = std::min(std::string(e) == "all" ? n_dev : std::atoi(e), n_dev);
                    // the scene is replicated: every further device gets its own context (kept across renders) and its own
                    // upload + BVH build, all of them at the same time on their own host threads
                    struct Job { int d; nrcu_ctx* ctx; bool created; int rc; std::string err; };
                    std::vector<Job> jobs;
                    for (int d = 0; d < n_dev && (int)jobs.size() + 1 < want; d++) {
                        if (d == dev) continue;
                        auto it = sh.extra.find(d);
                        jobs.push_back({d, it == sh.extra.end() ? nullptr : it->second, false, NRCU_OK, {}});
                    }
                    std::vector<std::thread> pool;
                    for (auto& j : jobs) pool.emplace_back([&j, &view]() {
                        if (!j.ctx) {
                            j.rc = nrcu_create(j.d, &j.ctx);
                            if (j.rc != NRCU_OK) { j.err = nrcu_last_error(nullptr); j.ctx = nullptr; return; }
                            j.created = true;
                        }
                        j.rc = nrcu_upload_scene(j.ctx, &view, NRCU_SCENE_MODE);
                        if (j.rc != NRCU_OK) j.err = nrcu_last_error(j.ctx);
                    });
                    for (auto& t : pool) t.join();
                    for (auto& j : jobs) {
                        if (j.created && j.ctx) sh.extra.emplace(j.d, j.ctx);
                        if (j.rc != NRCU_OK) { logger.warning("NRCuda: device " + std::to_string(j.d) + " skipped: " + j.err); continue; }   // a device that fails is left out, the others carry on
                        devs.push_back(j.ctx); dev_ids.push_back(j.d);
                    }
                }
                nrcu_render_params params{};
                if (const char* e = std::getenv("NRCU_SEED")) params.seed = std::strtoull(e, nullptr, 10);
                if (const char* e = std::getenv("NRCU_NEE")) if (std::atoi(e)) params.flags |= NRCU_FLAG_NEE;   // extension: same expectation, less noise
                if (const char* e = std::getenv("NRCU_ENV_IS")) if (std::atoi(e)) params.flags |= NRCU_FLAG_ENV_IS;   // extension: importance-sample the environment map
                if (const char* e = std::getenv("NRCU_GLASS_BRANCH")) params.glass_mode = std::atoi(e) ? NRCU_GLASS_BRANCH : NRCU_GLASS_STOCHASTIC;
                nrcu_stats st{};
                RGBA* pixels = sh.frameBuffer((size_t)w * h);   // the plugin owns the buffer, Screen::set copies it (RayCastRenderer.cpp:7-10)
                int rc;
                const char* prog = std::getenv("NRCU_PROGRESSIVE");
#if NRCU_PLUGIN_MODE == 3
                nrcu_mlt_params mp{};
                mp.seed = params.seed;
                mp.tone_map = NRCU_MLT_TONE_REFERENCE;           // Metropolis.cpp:118-123
                if (const char* e = std::getenv("NRCU_MLT_MUTATIONS_PER_PIXEL")) mp.mutations_per_pixel = (uint32_t)std::atoi(e);
                if (const char* e = std::getenv("NRCU_MLT_TONE")) mp.tone_map = (uint32_t)std::atoi(e);
                (void)prog;
                rc = nrcu_render_mlt(sh.ctx, &mp, reinterpret_cast<float*>(pixels), &st);
#else
                if (prog && std::atoi(prog) > 0 && devs.size() == 1) {
                    // publish intermediate frames: the GUI re-uploads whenever Screen::isUpdated() (ScreenView.cpp:168-173)
                    struct Pub { unsigned w, h; } pub{w, h};
                    rc = nrcu_render_progressive(sh.ctx, &params, (uint32_t)std::atoi(prog), reinterpret_cast<float*>(pixels),
                                                 [](void* u, const float* rgba, uint32_t, uint32_t) -> int {
                                                     auto* p = static_cast<Pub*>(u);
                                                     getServer().screen.set(reinterpret_cast<RGBA*>(const_cast<float*>(rgba)), (int)p->w, (int)p->h);
                                                     return 0;
                                                 }, &pub, &st);
                } else rc = nrcu_render_multi(devs.data(), (int)devs.size(), &params, reinterpret_cast<float*>(pixels), &st);
#endif
                if (rc != NRCU_OK) {
                    std::string why = nrcu_last_error(sh.ctx);   // the root context carries "device N: ..." for a failed peer
                    for (size_t g = 1; g < devs.size(); g++) {
                        const char* pe = nrcu_last_error(devs[g]);
                        if (pe && *pe && why.find(pe) == std::string::npos) why += " | device " + std::to_string(dev_ids[g]) + ": " + pe;
                    }
                    logger.error("NRCuda: " + why);
                    publishBlack(w, h);
                    return;
                }
                getServer().screen.set(pixels, (int)w, (int)h);
                double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                char buf[320];
                std::snprintf(buf, sizeof(buf),
                              "Done... %.3f s on %d GPU(s) (GPU %.1f ms, setup %.2f ms): %.2f Mpath-samples/s, %.1f Mrays/s, %u primitives, %u BVH4 nodes",
                              wall, (int)devs.size(), st.ms_total, st.ms_setup, st.paths / wall * 1e-6, st.rays / (st.ms_total * 1e-3) * 1e-6,
                              st.n_primitives, st.bvh_nodes);
                logger.log(buf);
            } catch (const std::exception& ex) {
                logger.error(std::string("NRCuda: ") + ex.what());
                publishBlack(w, h);
            } catch (...) {
                logger.error("NRCuda: unknown failure");
                publishBlack(w, h);
            }
        }
    };
}

#if NRCU_PLUGIN_MODE == 0
const static std::string description =
    "CUDA (sm_100a) Ray Cast Renderer.\n"
    "Same image as RayCast: Lambertian and Phong, one point light,\n"
    "triangle / sphere / plane, pinhole camera.";
REGISTER_RENDERER(CudaRayCast, description, NRCuda::Adapter);
#elif NRCU_PLUGIN_MODE == 1
const static std::string description =
    "CUDA (sm_100a) wavefront path tracer, SimplePathTracer semantics:\n"
    "Lambertian only, uniform hemisphere sampling, area lights hit by chance.";
REGISTER_RENDERER(CudaSimplePathTracer, description, NRCuda::Adapter);
#elif NRCU_PLUGIN_MODE == 3
const static std::string description =
    "CUDA (sm_100a) Metropolis light transport in primary sample space\n"
    "(one Markov chain per GPU thread over the path tracer's sampler).";
REGISTER_RENDERER(CudaMetropolisLightTransport, description, NRCuda::Adapter);
#else
const static std::string description =
    "CUDA (sm_100a) wavefront path tracer, AccPathTracer semantics:\n"
    "wide BVH, Lambertian / conductor / glass / microfacet materials, environment map.";
REGISTER_RENDERER(CudaAccPathTracer, description, NRCuda::Adapter);
#endif
