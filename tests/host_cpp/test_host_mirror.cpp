// C++ host mirror checks.  Without arguments: configuration surface only (no device needed).
// With "gpu": a short run through slime::CudaBackend on device 0 (prints a checksum the pytest side
// compares with the oracle).
#include <cassert>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../slime_mold_b200/host/slime_backend.hpp"

int main(int argc, char** argv)
{
    using namespace slime;
    Settings d;
    assert(d.agent_count == 10000000 && d.window_width == 1600 && d.window_height == 900);
    PresetManager pm = init_preset_manager();
    auto names = pm.get_preset_names();
    assert(names.size() == 8 && names[0] == "Default" && names[7] == "Mesh");
    assert(pm.get_preset("nope") == nullptr);
    const Settings& curls = pm.get_preset("Curls")->settings;
    assert(curls.agent_count == 3000000 && curls.pheromone_decay_factor == 75.0f && curls.agent_jitter == 5.0f);
    SimSizeUniform u = SimSizeUniform::create(1920, 1080, curls.pheromone_decay_factor, curls);
    unsigned char raw[56];
    std::memcpy(raw, &u, 56);
    // emit the packed uniforms of all presets as hex so the Python test can compare them byte for byte
    for (const auto& n : names) {
        const Settings& s = pm.get_preset(n)->settings;
        SimSizeUniform v = SimSizeUniform::create(1920, 1080, s.pheromone_decay_factor, s);
        std::memcpy(raw, &v, 56);
        std::printf("%s:", n.c_str());
        for (int i = 0; i < 56; ++i) std::printf("%02x", raw[i]);
        std::printf("\n");
    }
    if (argc > 1 && std::string(argv[1]) == "gpu") {
        Settings s = pm.get_preset("Waves")->settings;
        s.agent_count = 20000;
        CudaBackend be(256, 128, s);
        be.init_agents(5);
        be.step(10);
        auto a = be.read_agents();
        auto t = be.read_trail();
        unsigned long long h = 1469598103934665603ull;      // FNV-1a over the raw bits
        auto mixin = [&](const float* p, size_t n) {
            const unsigned char* b = reinterpret_cast<const unsigned char*>(p);
            for (size_t i = 0; i < 4 * n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        };
        mixin(a.data(), a.size());
        mixin(t.data(), t.size());
        std::printf("gpu_checksum:%016llx\n", h);
        try {
            SimSizeUniform bad = SimSizeUniform::create(64, 64, 1.0f, s);
            be.write_uniform(bad);
            std::printf("error_check:missing\n");
        } catch (const Error& e) {
            std::printf("error_check:ok %d\n", e.code);
        }
    } else if (argc > 1 && std::string(argv[1]) == "nodevice") {
        try {
            Settings s;
            s.agent_count = 16;
            CudaBackend be(64, 64, s);
            std::printf("nodevice:unexpected success\n");
        } catch (const Error& e) {
            std::printf("nodevice:%d %s\n", e.code, e.what());
        }
    }
    return 0;
}
