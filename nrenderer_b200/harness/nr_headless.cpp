// nr_headless — headless driver for NRenderer render components (the reference ships only a GUI).
//
// It reproduces what the GUI does when the user clicks "Render"
// (reference code/app/src/ui/views/SceneView.cpp:97-102): import files into an Asset
// (ScnImporter/ObjImporter::import, code/app/include/importer/Importer.hpp:16), build the Scene
// (SceneBuilder::build, code/app/src/asset/SceneBuilder.cpp:100-110), look the component up by
// name (ComponentFactory::createComponent, code/include/component/ComponentFactory.hpp:30-33) and
// call RenderComponent::exec (code/server/component/RenderComponent.cpp:5-9); the frame is read
// back from getServer().screen (code/server/server/Screen.cpp:54-66).
//
// It is linked against the reference's own importer/SceneBuilder/server sources (built by
// oracle/build_ref.py into oracle/_ref/), so it is both the parity oracle runner and the CPU
// baseline timer, and it can host the CUDA plugin exactly like the GUI would.  Scenes can also be
// read from / written to `.nrsc` flat-scene fixtures so that it works where /root/reference is absent.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "server/Server.hpp"
#include "component/RenderComponent.hpp"
#include "asset/Asset.hpp"
#include "asset/SceneBuilder.hpp"
#include "importer/ScnImporter.hpp"
#include "importer/ObjImporter.hpp"
#include "utilities/ImageLoader.hpp"

#include "../plugin/scene_bridge.hpp"
#include "ComponentManager.hpp"

// The importers call these after a successful parse to build OpenGL preview buffers
// (code/app/src/importer/ScnImporter.cpp:508-514, ObjImporter.cpp:396-398); headless: no-ops.
namespace NRenderer {
void Asset::genPreviewGlBuffersPerNode(NodeItem&) {}
void Asset::genPreviewGlBuffersPerLight(LightItem&) {}
void Asset::updateNodeGlDrawData(NodeItem&) {}
void Asset::updateLightGlDrawData(LightItem&) {}
}  // namespace NRenderer

using namespace NRenderer;

// Minimal PNG writer (8-bit RGB, zlib "stored" blocks: no compressor needed, any viewer reads it).
namespace {
uint32_t crc32_update(uint32_t c, const unsigned char* p, size_t n) {
    static uint32_t table[256];
    if (!table[1]) for (uint32_t i = 0; i < 256; i++) { uint32_t v = i; for (int k = 0; k < 8; k++) v = (v >> 1) ^ (0xEDB88320u & (0u - (v & 1u))); table[i] = v; }
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c;
}
void be32(std::vector<unsigned char>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s)); }
void png_chunk(std::ofstream& o, const char* tag, const std::vector<unsigned char>& body) {
    std::vector<unsigned char> c; be32(c, (uint32_t)body.size());
    c.insert(c.end(), tag, tag + 4); c.insert(c.end(), body.begin(), body.end());
    be32(c, crc32_update(0xFFFFFFFFu, c.data() + 4, c.size() - 4) ^ 0xFFFFFFFFu);
    o.write((const char*)c.data(), c.size());
}
void write_png_rgb8(std::ofstream& o, const std::vector<unsigned char>& rgb, unsigned w, unsigned h) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    o.write((const char*)sig, 8);
    std::vector<unsigned char> ihdr; be32(ihdr, w); be32(ihdr, h);
    for (unsigned char b : {8, 2, 0, 0, 0}) ihdr.push_back(b);          // 8 bit, colour type 2 (RGB)
    png_chunk(o, "IHDR", ihdr);
    std::vector<unsigned char> raw; raw.reserve((size_t)h * (3 * (size_t)w + 1));
    for (unsigned y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgb.begin() + (size_t)y * 3 * w, rgb.begin() + (size_t)(y + 1) * 3 * w); }
    std::vector<unsigned char> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < raw.size(); i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back((unsigned char)(n & 0xFF)); z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xFF)); z.push_back((unsigned char)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        if (raw.empty()) break;
    }
    be32(z, (b << 16) | a);
    png_chunk(o, "IDAT", z);
    png_chunk(o, "IEND", {});
}
}  // namespace

static void usage() {
    std::fprintf(stderr,
        "usage: nr_headless [--scn F|--obj F]... | --flat F.nrsc\n"
        "         [--mesh-material K] [--texture IMG] [--env-map TEXIDX]\n"
        "         [--w W --h H --depth D --spp S --aspect A --ambient R G B]\n"
        "         [--dump-flat OUT.nrsc]\n"
        "         [--plugin LIB.so]... [--plugin-dir DIR] [--manager]\n"
        "         [--component NAME --out FRAME.f32|.ppm|.png|.pfm] [--warmup W] [--repeat N] [--list]\n"
        "  --plugin-dir DIR  load every *.so of DIR like the GUI's ComponentManager::init (ComponentManager.cpp:15-30)\n"
        "  --manager         run the component the way the GUI does: ComponentManager::exec on a detached thread,\n"
        "                    polling getState() and Screen::isUpdated() (ComponentProgressView.cpp, ScreenView.cpp:168-173)\n");
}

int main(int argc, char** argv) {
    std::vector<std::pair<std::string, std::string>> imports;
    std::vector<std::string> plugins, textures, plugin_dirs;
    std::string flat_in, flat_out, component, out;
    int mesh_material = -1, env_map = -1, repeat = 1, warmup = 0;
    bool list = false, via_manager = false;
    long w = -1, h = -1, depth = -1, spp = -1;
    float aspect = -1.f;
    bool have_ambient = false;
    float ambient[3] = {0, 0, 0};
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) { usage(); std::exit(2); }
            return argv[++i];
        };
        if (a == "--scn") imports.push_back({"scn", next()});
        else if (a == "--obj") imports.push_back({"obj", next()});
        else if (a == "--flat") flat_in = next();
        else if (a == "--dump-flat") flat_out = next();
        else if (a == "--plugin") plugins.push_back(next());
        else if (a == "--plugin-dir") plugin_dirs.push_back(next());
        else if (a == "--manager") via_manager = true;
        else if (a == "--warmup") warmup = std::atoi(next().c_str());
        else if (a == "--component") component = next();
        else if (a == "--out") out = next();
        else if (a == "--mesh-material") mesh_material = std::atoi(next().c_str());
        else if (a == "--texture") textures.push_back(next());
        else if (a == "--env-map") env_map = std::atoi(next().c_str());
        else if (a == "--w") w = std::atol(next().c_str());
        else if (a == "--h") h = std::atol(next().c_str());
        else if (a == "--depth") depth = std::atol(next().c_str());
        else if (a == "--spp") spp = std::atol(next().c_str());
        else if (a == "--aspect") aspect = (float)std::atof(next().c_str());
        else if (a == "--repeat") repeat = std::atoi(next().c_str());
        else if (a == "--ambient") { have_ambient = true; for (int k = 0; k < 3; k++) ambient[k] = (float)std::atof(next().c_str()); }
        else if (a == "--list") list = true;
        else { usage(); return 2; }
    }

    SharedScene scene;
    if (!flat_in.empty()) {
        nrb200::FlatScene f;
        try { f = nrb200::load_flat_scene(flat_in); }
        catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
        scene = nrb200::unflatten(f);
    } else if (!imports.empty()) {
        Asset asset;
        for (auto& [kind, path] : imports) {
            bool ok; std::string err;
            if (kind == "scn") { ScnImporter imp; ok = imp.import(asset, path); err = imp.getErrorInfo(); }
            else { ObjImporter imp; ok = imp.import(asset, path); err = imp.getErrorInfo(); }
            if (!ok) { std::fprintf(stderr, "import of %s failed: %s\n", path.c_str(), err.c_str()); return 1; }
        }
        if (mesh_material >= 0)
            for (auto& m : asset.meshes) m->material.setIndex((unsigned)mesh_material);
        RenderSettings rs; AmbientSettings as; Camera cam;   // GUI defaults (RenderSettingsManager.hpp:9-45)
        SceneBuilder sb{asset, rs, as, cam};
        scene = sb.build();
        if (!scene) { std::fprintf(stderr, "SceneBuilder::build failed (a node has no material)\n"); return 1; }
    }
    if (scene) {
        if (w > 0) scene->renderOption.width = (unsigned)w;
        if (h > 0) scene->renderOption.height = (unsigned)h;
        if (depth >= 0) scene->renderOption.depth = (unsigned)depth;
        if (spp > 0) scene->renderOption.samplesPerPixel = (unsigned)spp;
        if (aspect > 0) scene->camera.aspect = aspect;
        if (have_ambient) scene->ambient.constant = {ambient[0], ambient[1], ambient[2]};
        for (auto& path : textures) {   // TextureImporter semantics minus the GL upload (TextureImporter.cpp:7-21)
            ImageLoader loader;
            Image* img = loader.load(path);
            if (!img || !img->data) { std::fprintf(stderr, "cannot load texture %s\n", path.c_str()); return 1; }
            Texture t; t.width = img->width; t.height = img->height;
            t.rgba = new RGBA[(size_t)t.width * t.height];
            std::memcpy((void*)t.rgba, img->data, sizeof(float) * 4 * (size_t)t.width * t.height);
            scene->textures.push_back(std::move(t));
            delete img;
        }
        if (env_map >= 0) {
            scene->ambient.type = Ambient::Type::ENVIROMENT_MAP;
            scene->ambient.environmentMap = Handle{(unsigned)env_map};
        }
    }
    if (!flat_out.empty()) {
        if (!scene) { std::fprintf(stderr, "--dump-flat needs a scene\n"); return 2; }
        nrb200::save_flat_scene(nrb200::flatten(*scene), flat_out);
    }

    for (auto& p : plugins) {
        // Same effect as the GUI's LoadLibrary loop (code/app/src/manager/ComponentManager.cpp:15-30):
        // the library's static ComponentRegister object registers the component.  RTLD_LOCAL: every
        // plugin defines a struct named ComponentRegister (Component.hpp:23-32); with global binding the
        // second plugin would run the first one's constructor.  getServer() is still shared because all
        // plugins depend on the one libNRServer.so.
        if (!dlopen(p.c_str(), RTLD_NOW | RTLD_LOCAL)) { std::fprintf(stderr, "dlopen %s: %s\n", p.c_str(), dlerror()); return 1; }
    }
    ComponentManager manager;
    for (auto& d : plugin_dirs) {
        manager.init(d + "/*.so");
        for (auto& e : manager.getLoadErrors()) std::fprintf(stderr, "plugin-dir %s: %s\n", d.c_str(), e.c_str());
    }
    if (list) {
        for (auto& ci : getServer().componentFactory.getComponentsInfo("Render"))
            std::printf("%s\n", ci.name.c_str());
    }
    if (component.empty()) return 0;
    if (!scene) { std::fprintf(stderr, "no scene\n"); return 2; }

    double best = 1e300, total = 0;
    unsigned screen_updates = 0;
    ComponentInfo info;
    for (auto& ci : getServer().componentFactory.getComponentsInfo("Render")) if (ci.name == component) info = ci;
    for (int r = -warmup; r < repeat; r++) {
        // The reference components mutate the Scene in place, so every run gets a fresh copy
        // (the GUI builds a new Scene per click, SceneView.cpp:100-101).
        SharedScene run_scene = std::make_shared<Scene>(*scene);
        double s;
        if (via_manager) {
            auto t0 = std::chrono::steady_clock::now();
            if (!manager.exec<RenderComponent>(info, run_scene)) { std::fprintf(stderr, "component %s is not registered\n", component.c_str()); return 1; }
            // what the GUI's frame loop does while the worker runs: look at the state, re-read the screen when it changed
            while (manager.getState() != ComponentManager::State::FINISH) {
                if (getServer().screen.isUpdated()) { (void)getServer().screen.getPixels(); if (r >= 0) screen_updates++; }
                std::this_thread::sleep_for(std::chrono::microseconds(200));
            }
            s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (getServer().screen.isUpdated()) { (void)getServer().screen.getPixels(); if (r >= 0) screen_updates++; }
            manager.finish();
        } else {
            auto comp = getServer().componentFactory.createComponent<RenderComponent>("Render", component);
            if (!comp) { std::fprintf(stderr, "component %s is not registered\n", component.c_str()); return 1; }
            auto t0 = std::chrono::steady_clock::now();
            comp->exec([] {}, [] {}, run_scene);
            s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (r < 0) continue;   // warm-up runs are not timed
        best = std::min(best, s); total += s;
    }
    auto& screen = getServer().screen;
    unsigned sw = screen.getWidth(), sh = screen.getHeight();
    if (!out.empty()) {
        // The reference has no image export at all.  By extension:
        //   .ppm / .png  8-bit RGB of the published frame (already sqrt-gamma'd and clamped by Screen::set), row 0 = top
        //   .pfm  RGB fp32, little endian, rows bottom-to-top as the format requires
        //   else  raw RGBA fp32 exactly as getServer().screen holds it (what the parity tests read)
        const RGBA* px = screen.getPixels();
        auto ends_with = [&](const char* ext) { std::string e(ext); return out.size() >= e.size() && out.compare(out.size() - e.size(), e.size(), e) == 0; };
        std::ofstream o(out, std::ios::binary);
        if (ends_with(".ppm") || ends_with(".png")) {
            std::vector<unsigned char> rgb(3 * (size_t)sw * sh);
            for (size_t i = 0; i < (size_t)sw * sh; i++) for (int c = 0; c < 3; c++) {
                float v = px[i][c]; v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
                rgb[3 * i + c] = (unsigned char)(v * 255.f + 0.5f);
            }
            if (ends_with(".png")) write_png_rgb8(o, rgb, sw, sh);
            else { o << "P6\n" << sw << " " << sh << "\n255\n"; o.write((const char*)rgb.data(), rgb.size()); }
        } else if (ends_with(".pfm")) {
            o << "PF\n" << sw << " " << sh << "\n-1.0\n";
            std::vector<float> row(3 * (size_t)sw);
            for (unsigned y = sh; y-- > 0;) {
                for (unsigned x = 0; x < sw; x++) for (int c = 0; c < 3; c++) row[3 * x + c] = px[(size_t)y * sw + x][c];
                o.write((const char*)row.data(), row.size() * sizeof(float));
            }
        } else {
            o.write((const char*)px, sizeof(float) * 4 * (size_t)sw * sh);
        }
    }
    std::string last_log, last_error;
    unsigned n_errors = 0, n_warnings = 0;
    {
        auto logs = getServer().logger.get();
        for (unsigned i = 0; i < logs.nums; i++) {
            last_log = logs.msgs[i].message;
            if (logs.msgs[i].type == Logger::LogType::ERROR) { n_errors++; last_error = logs.msgs[i].message; }
            if (logs.msgs[i].type == Logger::LogType::WARNING) n_warnings++;
        }
    }
    for (auto* str : {&last_log, &last_error}) for (auto& c : *str) if (c == '"' || c == '\n' || c == '\\') c = ' ';
    std::printf("{\"component\": \"%s\", \"width\": %u, \"height\": %u, \"spp\": %u, \"depth\": %u, "
                "\"seconds\": %.6f, \"seconds_mean\": %.6f, \"repeat\": %d, \"warmup\": %d, \"host_threads\": %u, \"via\": \"%s\", "
                "\"screen_updates\": %u, \"errors\": %u, \"warnings\": %u, \"last_error\": \"%s\", \"last_log\": \"%s\"}\n",
                component.c_str(), sw, sh, scene->renderOption.samplesPerPixel, scene->renderOption.depth,
                best, total / repeat, repeat, warmup, std::thread::hardware_concurrency(),
                via_manager ? "ComponentManager::exec (detached thread)" : "RenderComponent::exec",
                screen_updates, n_errors, n_warnings, last_error.c_str(), last_log.c_str());
    return 0;
}
