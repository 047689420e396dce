// ComponentManager.hpp — a Windows-free NRenderer::ComponentManager.
//
// Same public interface and state machine as the reference's
//   code/app/include/manager/ComponentManager.hpp:15-70   (class, State {IDLING, READY, RUNNING, FINISH}, exec<>)
//   code/app/src/manager/ComponentManager.cpp:15-58       (init = _findfirst + LoadLibrary over "<dir>\*.dll")
// so that the reference's ImGui app (Manager::componentManager, ComponentProgressView.cpp:14-41, SceneView.cpp:97-102)
// could host the CUDA plugins unchanged on Linux.  What differs is the mechanism: POSIX opendir/dlopen instead
// of the Win32 loader, atomics + a condition variable instead of plain fields written from the worker thread
// (the reference polls `state` from the UI thread without synchronisation), and `wait()` for headless callers.
//
// Each plugin is opened RTLD_LOCAL: REGISTER_COMPONENT defines a struct of the same name (ComponentRegister) in
// every plugin (Component.hpp:23-32); with global symbol binding the second library would run the first one's
// constructor.  getServer() stays shared because every plugin links the one libNRServer.so.
#pragma once
#include <dirent.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "component/RenderComponent.hpp"
#include "server/Server.hpp"

namespace NRenderer
{
    class ComponentManager
    {
    public:
        enum class State { IDLING, READY, RUNNING, FINISH };

    private:
        struct Shared {   // outlives the manager if a detached worker is still running
            std::atomic<State> state{State::IDLING};
            std::mutex mtx;
            std::condition_variable cv;
            std::chrono::system_clock::time_point lastStartTime{}, lastEndTime{};
        };
        std::shared_ptr<Shared> sh = std::make_shared<Shared>();
        std::vector<void*> loadedLibraries;
        std::vector<std::string> loadErrors;
        ComponentInfo activeComponent;

    public:
        ComponentManager() = default;
        ComponentManager(const ComponentManager&) = delete;
        ~ComponentManager() {
            // Libraries stay mapped: a detached render thread may still be inside one, and unloading would run the
            // plugin's ~ComponentRegister while the factory is being torn down at exit (the reference FreeLibrary()s
            // here and relies on process exit order).
        }

        // `path` is what the reference passes: "<dir>\*.dll" (or "<dir>/*.so", or just a directory).  Every shared
        // object in the directory is loaded; whatever registers itself with the ComponentFactory becomes available.
        void init(const std::string& path) {
            std::string dir = path;
            const size_t star = dir.find('*');
            if (star != std::string::npos) dir = dir.substr(0, star);
            while (!dir.empty() && (dir.back() == '/' || dir.back() == '\\')) dir.pop_back();
            if (dir.empty()) dir = ".";
            DIR* d = opendir(dir.c_str());
            if (!d) return;   // like the reference: a missing directory is not an error
            std::vector<std::string> names;
            while (dirent* e = readdir(d)) {
                std::string n = e->d_name;
                if (n.size() > 3 && n.compare(n.size() - 3, 3, ".so") == 0) names.push_back(n);
            }
            closedir(d);
            std::sort(names.begin(), names.end());   // readdir order is arbitrary; keep registration order stable
            for (auto& n : names) {
                void* h = dlopen((dir + "/" + n).c_str(), RTLD_NOW | RTLD_LOCAL);
                if (h) loadedLibraries.push_back(h);
                else loadErrors.push_back(n + ": " + dlerror());
            }
        }
        const std::vector<std::string>& getLoadErrors() const { return loadErrors; }
        size_t getLoadedCount() const { return loadedLibraries.size(); }

        ComponentInfo getActiveComponentInfo() const { return activeComponent; }

        // createComponent -> READY -> detached thread running Interface::exec(onStart, onFinish, args...)
        // (onStart: RUNNING + start time, onFinish: FINISH + end time), ComponentManager.hpp:41-64.
        // Returns false (state stays IDLING) when the component is not registered - the reference would
        // dereference the null shared_ptr on the worker thread.
        template <typename Interface, typename... Args>
        bool exec(const ComponentInfo& componentInfo, Args... args) {
            auto component = getServer().componentFactory.createComponent<Interface>(componentInfo.type, componentInfo.name);
            if (!component) return false;
            activeComponent = componentInfo;
            sh->state = State::READY;
            auto s = sh;
            try {
                std::thread t([s, component, args...]() mutable {
                    auto done = [s]() {
                        {
                            std::lock_guard<std::mutex> lk(s->mtx);
                            s->lastEndTime = std::chrono::system_clock::now();
                            s->state = State::FINISH;
                        }
                        s->cv.notify_all();
                    };
                    try {
                        component->exec(
                            [s]() { std::lock_guard<std::mutex> lk(s->mtx); s->lastStartTime = std::chrono::system_clock::now(); s->state = State::RUNNING; },
                            done, args...);
                    } catch (...) {   // a throwing component must not take the host down (the reference would terminate)
                        std::cerr << "Unexpected termination" << std::endl;
                        done();
                    }
                });
                t.detach();
            } catch (const std::exception& e) {
                std::cerr << "Unexpected termination" << std::endl << e.what() << std::endl;
                sh->state = State::IDLING;
                return false;
            }
            return true;
        }

        void finish() { sh->state = State::IDLING; }
        State getState() const { return sh->state.load(); }
        std::chrono::duration<double> getLastExecTime() const {
            std::lock_guard<std::mutex> lk(sh->mtx);
            return sh->lastEndTime - sh->lastStartTime;
        }

        // Headless hosts: block until the running component reports FINISH (the GUI polls getState() per frame
        // instead).  Returns false on timeout.
        bool wait(double timeout_seconds = 1e9) {
            std::unique_lock<std::mutex> lk(sh->mtx);
            return sh->cv.wait_for(lk, std::chrono::duration<double>(timeout_seconds),
                                   [&] { State st = sh->state.load(); return st == State::FINISH || st == State::IDLING; });
        }
    };
}  // namespace NRenderer
