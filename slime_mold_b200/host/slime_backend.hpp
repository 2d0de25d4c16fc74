// slime_backend.hpp -- C++ host-side mirror of the reference's Rust surface for the step loop.
//
// The reference is compiled code (Rust); rustc/cargo are not available in this image, so the host
// side above the C ABI is written in C++ with the reference's own names and argument meaning:
//
//   slime::Settings            <- src/settings.rs:29-70   (same fields, same defaults)
//   slime::Preset / PresetManager / init_preset_manager  <- src/presets.rs:6-155
//   slime::SimSizeUniform      <- src/main.rs:29-67       (== C `sm_params`, 56 bytes)
//   slime::CudaBackend         <- the compute halves of src/pipeline_manager.rs:7-75 and
//                                 src/bind_group_manager.rs:5-65, plus the buffer operations
//                                 src/main.rs performs on them (line refs on each method)
//
// Error behaviour: the reference unwrap()s / panics on every failure (src/main.rs:169,182,207,...);
// here every failing C-ABI call throws slime::Error carrying sm_last_error().  There is no CPU
// fallback.  Header only; link with libslime_b200.so.
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/slime_b200.h"

namespace slime {

// ---- settings.rs:4-27 -----------------------------------------------------------------------
constexpr uint32_t DEFAULT_WIDTH = 1600;
constexpr uint32_t DEFAULT_HEIGHT = 900;
constexpr bool DEFAULT_IS_FULLSCREEN = false;
constexpr size_t AGENT_COUNT = 10'000'000;
constexpr float AGENT_SPEED_MIN = 30.0f;
constexpr float AGENT_SPEED_MAX = 50.0f;
constexpr float AGENT_TURN_SPEED = 0.43f;
constexpr float AGENT_POSSIBLE_STARTING_HEADINGS_START = 0.0f;
constexpr float AGENT_POSSIBLE_STARTING_HEADINGS_END = 360.0f;
constexpr float DEPOSITION_AMOUNT = 1.0f;
constexpr float AGENT_JITTER = 0.0f;
constexpr float AGENT_SENSOR_ANGLE = 0.3f;
constexpr float AGENT_SENSOR_DISTANCE = 20.0f;
constexpr float DECAY_FACTOR = 10.0f;
constexpr float DIFFUSION_RATE = 1.0f;
constexpr float BLUR_RADIUS = 2.0f;
constexpr float BLUR_SIGMA = 1.0f;

// settings.rs:29-47 (field order kept)
struct Settings {
    size_t agent_count = AGENT_COUNT;
    float agent_jitter = AGENT_JITTER;
    std::pair<float, float> agent_possible_starting_headings{AGENT_POSSIBLE_STARTING_HEADINGS_START,
                                                             AGENT_POSSIBLE_STARTING_HEADINGS_END};
    float agent_speed_max = AGENT_SPEED_MAX;
    float agent_speed_min = AGENT_SPEED_MIN;
    float agent_turn_speed = AGENT_TURN_SPEED;
    float pheromone_decay_factor = DECAY_FACTOR;
    float pheromone_diffusion_rate = DIFFUSION_RATE;
    float pheromone_deposition_amount = DEPOSITION_AMOUNT;
    bool window_fullscreen = DEFAULT_IS_FULLSCREEN;
    uint32_t window_height = DEFAULT_HEIGHT;
    uint32_t window_width = DEFAULT_WIDTH;
    float agent_sensor_angle = AGENT_SENSOR_ANGLE;
    float agent_sensor_distance = AGENT_SENSOR_DISTANCE;
    float blur_radius = BLUR_RADIUS;
    float blur_sigma = BLUR_SIGMA;
};

// ---- presets.rs ---------------------------------------------------------------------------------
struct Preset {
    std::string name;
    Settings settings;
};

class PresetManager {
public:
    void add_preset(Preset p) { presets_.push_back(std::move(p)); }
    const Preset* get_preset(const std::string& name) const
    {
        for (const auto& p : presets_)
            if (p.name == name) return &p;
        return nullptr;
    }
    std::vector<std::string> get_preset_names() const
    {
        std::vector<std::string> n;
        for (const auto& p : presets_) n.push_back(p.name);
        return n;
    }

private:
    std::vector<Preset> presets_;
};

namespace detail {
inline Settings preset(float jitter, float smin, float smax, float turn, float sangle, float sdist, float dep, float decay,
                       float diffusion, size_t agents = AGENT_COUNT)
{
    Settings s;
    s.agent_count = agents;
    s.agent_jitter = jitter;
    s.agent_speed_min = smin;
    s.agent_speed_max = smax;
    s.agent_turn_speed = turn;
    s.agent_sensor_angle = sangle;
    s.agent_sensor_distance = sdist;
    s.pheromone_deposition_amount = dep;
    s.pheromone_decay_factor = decay;
    s.pheromone_diffusion_rate = diffusion;
    return s;
}
}  // namespace detail

// presets.rs:45-155
inline PresetManager init_preset_manager()
{
    PresetManager pm;
    pm.add_preset({"Default", Settings{}});
    pm.add_preset({"Sponge", detail::preset(0.0f, 20.0f, 30.0f, 0.43f, 0.3f, 20.0f, 1.0f, 1.0f, 1.0f)});
    pm.add_preset({"Firecracker Trees", detail::preset(0.1f, 60.0f, 60.0f, 1.47f, 0.3f, 20.0f, 1.0f, 10.0f, 1.0f)});
    pm.add_preset({"Threads", detail::preset(0.0f, 70.0f, 80.0f, 0.02f, 0.3f, 20.0f, 1.0f, 10.0f, 0.1f)});
    pm.add_preset({"Curls", detail::preset(5.0f, 70.0f, 80.0f, 0.05f, 0.3f, 20.0f, 1.0f, 75.0f, 0.1f, 3'000'000)});
    pm.add_preset({"Waves", detail::preset(1.0f, 30.0f, 50.0f, 6.0f, 0.3f, 20.0f, 1.0f, 10.0f, 0.1f)});
    pm.add_preset({"Snake", detail::preset(3.0f, 100.0f, 120.0f, 0.37f, 1.34f, 225.0f, 1.0f, 10.0f, 1.0f)});
    pm.add_preset({"Mesh", detail::preset(3.0f, 100.0f, 120.0f, 6.0f, 1.57f, 225.0f, 1.0f, 10.0f, 1.0f)});
    return pm;
}

// ---- main.rs:29-67 ------------------------------------------------------------------------------
struct SimSizeUniform : sm_params {
    static SimSizeUniform create(uint32_t width, uint32_t height, float decay_factor, const Settings& s)
    {
        SimSizeUniform u{};
        u.width = width;
        u.height = height;
        u.decay_factor = decay_factor;
        u.agent_jitter = s.agent_jitter;
        u.agent_speed_min = s.agent_speed_min;
        u.agent_speed_max = s.agent_speed_max;
        u.agent_turn_speed = s.agent_turn_speed;
        u.agent_sensor_angle = s.agent_sensor_angle;
        u.agent_sensor_distance = s.agent_sensor_distance;
        u.diffusion_rate = s.pheromone_diffusion_rate;
        u.pheromone_deposition_amount = s.pheromone_deposition_amount;
        u.blur_radius = s.blur_radius;
        u.blur_sigma = s.blur_sigma;
        u._pad = 0;
        return u;
    }
};
static_assert(sizeof(SimSizeUniform) == 56, "SimSizeUniform must stay the reference's 56-byte block");

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// ---- the compute backend -------------------------------------------------------------------------
class CudaBackend {
public:
    // PipelineManager::new + BindGroupManager::new + buffer creation (main.rs:263-293, 326-368)
    CudaBackend(uint32_t width, uint32_t height, const Settings& settings, int device = 0, int rank = 0, int world_size = 1,
                uint32_t flags = 0)
        : width_(width), height_(height), settings_(settings)
    {
        sm_config cfg{};
        cfg.width = width;
        cfg.height = height;
        cfg.agent_count = settings.agent_count;
        cfg.device = device;
        cfg.rank = rank;
        cfg.world_size = world_size;
        cfg.flags = flags;
        check(sm_create(&h_, &cfg));
        update_settings(settings);
    }
    ~CudaBackend() { if (h_) sm_destroy(h_); }
    CudaBackend(const CudaBackend&) = delete;
    CudaBackend& operator=(const CudaBackend&) = delete;

    // update_settings(), main.rs:83-99: rebuild the uniform and write it
    void update_settings(const Settings& s)
    {
        settings_ = s;
        write_uniform(SimSizeUniform::create(width_, height_, s.pheromone_decay_factor, s));
    }
    // queue.write_buffer(&sim_size_buffer, 0, bytes_of(&uniform)), main.rs:98
    void write_uniform(const SimSizeUniform& u) { check(sm_set_params(h_, &u)); }

    void init_agents(uint64_t seed) { check(sm_init_agents(h_, seed)); }                       // main.rs:269-282 (seeded)
    void write_agents(const float* xyas, uint64_t first, uint64_t n) { check(sm_upload_agents(h_, xyas, first, n)); }   // :142, :994
    std::vector<float> read_agents()                                                            // :121-131
    {
        std::vector<float> a(4 * sm_agent_count(h_));
        check(sm_download_agents(h_, a.data(), 0, sm_agent_count(h_), nullptr));
        return a;
    }
    void reassign_agent_speeds(uint64_t seed) { check(sm_reassign_speeds(h_, seed)); }          // main.rs:101-145
    void set_agent_count(uint64_t n, uint64_t seed)                                             // main.rs:682-791
    {
        check(sm_set_agent_count(h_, n, seed));
        settings_.agent_count = n;
    }
    void clear_trail() { check(sm_clear_trail(h_)); }                                           // main.rs:909-913
    std::vector<float> read_trail()
    {
        std::vector<float> t((size_t)width_ * height_);
        check(sm_download_trail(h_, t.data(), 0, 0, width_, height_, width_));
        return t;
    }
    void write_trail(const float* t) { check(sm_upload_trail(h_, t, 0, 0, width_, height_, width_)); }
    void resize(uint32_t w, uint32_t h)                                                         // main.rs:954-1015
    {
        check(sm_resize(h_, w, h));
        width_ = w;
        height_ = h;
    }
    // one frame: agents -> decay -> diffuse (main.rs:1163-1235)
    // display pass: main.rs:1202-1217 / display.wgsl; `lut768` = LutData.red ++ green ++ blue (main.rs:330-334)
    void set_lut(const uint8_t* lut768) { check(sm_set_lut(h_, lut768)); }
    std::vector<uint8_t> render(uint32_t tex_width, uint32_t tex_height)
    {
        std::vector<uint8_t> rgba((size_t)tex_width * tex_height * 4);
        check(sm_render_rgba8(h_, tex_width, tex_height, rgba.data()));
        return rgba;
    }
    void save_snapshot(const std::string& path) { check(sm_save_snapshot(h_, path.c_str())); }
    void load_snapshot(const std::string& path) { check(sm_load_snapshot(h_, path.c_str())); }
    void step(uint32_t n_steps = 1) { check(sm_step(h_, n_steps)); }
    void diffuse_only(uint32_t n) { check(sm_diffuse_only(h_, n)); }
    void sync() { check(sm_sync(h_)); }
    sm_trail_stats trail_statistics()
    {
        sm_trail_stats s{};
        check(sm_trail_statistics(h_, &s));
        return s;
    }
    uint64_t agent_count() const { return sm_agent_count(h_); }
    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    const Settings& settings() const { return settings_; }
    sm_engine* handle() { return h_; }

private:
    static void check(int rc)
    {
        if (rc != SM_OK) throw Error(rc, sm_last_error());
    }
    sm_engine* h_ = nullptr;
    uint32_t width_, height_;
    Settings settings_;
};

}  // namespace slime
