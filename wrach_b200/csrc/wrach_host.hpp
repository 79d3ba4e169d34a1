// wrach_host.hpp — C++ restatement of Wrach's Rust host side for the physics step, with the
// reference's names.  See include/wrach_host.h for the file:line map.  Header-only data model;
// the systems that talk to the CUDA worker are in wrach_host.cpp.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <variant>
#include <vector>

#include "../../include/wrach_host.h"

namespace wrach::host {

struct Vec2 { float x = 0, y = 0; };
struct Vec4 { float x = 0, y = 0, z = 0, w = 0; };
struct IVec2 {
    int32_t x = 0, y = 0;
    bool operator==(const IVec2 &o) const { return x == o.x && y == o.y; }
};
struct UVec2 { uint32_t x = 0, y = 0; };
using SpatialBinCoord = IVec2;  // spatial_bin.rs:7-8

struct Particle { Vec2 position, velocity; };  // state.rs:41-46

struct WrachConfig {  // config_app.rs:10-36
    uint16_t dimensions[2] = {480, 352};
    bool boundaries_as_dimensions = false;
    uint16_t cell_size = 3;  // shaders/shared/src/lib.rs:34
};

using ShaderWorldSettings = wrach_world_settings;  // config_shader.rs:15-29

// f32::div_euclid followed by `as i32` (saturating, NaN -> 0): spatial_bin.rs:60-61
inline int32_t div_euclid_as_i32(float a, float b) {
    float q = std::trunc(a / b);
    if (std::fmod(a, b) < 0.0f) q = b > 0.0f ? q - 1.0f : q + 1.0f;
    if (q != q) return 0;
    if (q >= 2147483648.0f) return INT32_MAX;
    if (q <= -2147483648.0f) return INT32_MIN;
    return (int32_t)q;
}

// Storage of the arrays that travel to and from the GPU.  Large ones (>= 32 MB) come from
// cudaMallocHost: page-locked from the start and placed on the NUMA node next to the GPU -- copies run
// at 55 GB/s against 52 for registered malloc memory and 17 for pageable memory (B200 box).  Small ones,
// and everything on a machine without a CUDA device, are ordinary heap memory.
void *host_array_alloc(size_t bytes);
void host_array_free(void *p, size_t bytes);
template <typename T>
struct HostArrayAlloc {
    using value_type = T;
    HostArrayAlloc() = default;
    template <typename U>
    HostArrayAlloc(const HostArrayAlloc<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(host_array_alloc(n * sizeof(T))); }
    void deallocate(T *p, size_t n) { host_array_free(p, n * sizeof(T)); }
    template <typename U>
    bool operator==(const HostArrayAlloc<U> &) const { return true; }
    template <typename U>
    bool operator!=(const HostArrayAlloc<U> &) const { return false; }
};
template <typename T>
using HostArray = std::vector<T, HostArrayAlloc<T>>;

struct PackedData {  // spatial_bin.rs:21-34
    HostArray<uint32_t> indices;
    HostArray<Vec2> positions, velocities;
};

struct ParticleData {  // particle_store.rs:31-38
    std::vector<Vec2> positions, velocities;
};

struct CoordHash {
    size_t operator()(const IVec2 &c) const {
        uint64_t v = ((uint64_t)(uint32_t)c.x << 32) | (uint32_t)c.y;
        v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33;
        return (size_t)v;
    }
};

class ParticleStore;

class SpatialBin {  // spatial_bin.rs:10-149
  public:
    uint16_t cell_size = 3;
    UVec2 grid_dimensions;
    Vec4 viewport;
    // Strip workers (new, not in the reference): pack only the cell columns [col_begin, col_end) of
    // the active grid, row-major over those columns.  The default covers every column.
    uint32_t col_begin = 0, col_end = 0xFFFFFFFFu;
    SpatialBin() = default;
    SpatialBin(uint16_t cell_size_, Vec4 viewport_) : cell_size(cell_size_), viewport(viewport_) { update_grid_size(); }

    SpatialBinCoord get_cell_coord(Vec2 position) const {  // :48-64
        const float cs = (float)cell_size;
        return {div_euclid_as_i32(position.x, cs), div_euclid_as_i32(position.y, cs)};
    }
    // :68-89 without materialising the list: first cell + inclusive dimensions, row-major order
    void get_active_cells(SpatialBinCoord &bottom_left, UVec2 &grid) const {
        bottom_left = get_cell_coord({viewport.x, viewport.y});
        const SpatialBinCoord top_right = get_cell_coord({viewport.z, viewport.w});
        grid.y = top_right.y >= bottom_left.y ? (uint32_t)(top_right.y - bottom_left.y + 1) : 0u;
        grid.x = (grid.y && top_right.x >= bottom_left.x) ? (uint32_t)(top_right.x - bottom_left.x + 1) : 0u;
    }
    PackedData create_packed_data(const ParticleStore &store) const;  // :103-149

    void update_grid_size() {  // :92-95
        SpatialBinCoord bl;
        get_active_cells(bl, grid_dimensions);
    }
};

// particle_store.rs:14-133.  The reference keeps a HashMap<cell, ParticleData>; insertion order
// inside a cell is what create_packed_data emits.  Same here, except that bulk inserts are first
// appended to a flat log and bucketed lazily, so adding millions of particles stays linear.
class ParticleStore {
  public:
    std::unordered_map<SpatialBinCoord, ParticleData, CoordHash> hashmap;
    SpatialBin spatial_bin;
    uint32_t particles_in_frame_count = 0;
    std::vector<SpatialBinCoord> cells_to_read_from_gpu;  // declared by the reference, unused there too

    ParticleStore() = default;
    ParticleStore(uint16_t cell_size, Vec4 viewport) : spatial_bin(cell_size, viewport) {}

    void add_particle(const Particle &p) {  // :54-59
        log_.push_back(p);
    }
    void add_particles_to_cell(SpatialBinCoord cell, ParticleData particles) {  // :62-64
        flush_log();
        hashmap[cell] = std::move(particles);
    }
    void remove(SpatialBinCoord cell) {  // :67-69
        flush_log();
        hashmap.erase(cell);
    }
    // particle_store.rs:76-85 -- a commented-out stub in the reference ("Take `PackedData` from the
    // GPU and write back into the store"), completed here: the i-th active cell,
    // row-major from the viewport's first cell, owns the slots [indices[i+1], indices[i+2]) of what
    // `tick` read back; its bucket is replaced by them (their packed order becomes the bucket's
    // insertion order).  Cells outside the viewport keep what they hold.  Returns false when
    // `update` does not have the layout of the current viewport.
    bool update_from_gpu(const PackedData &update) {
        flush_log();
        SpatialBinCoord bl;
        UVec2 grid;
        spatial_bin.get_active_cells(bl, grid);
        const uint32_t c0 = spatial_bin.col_begin < grid.x ? spatial_bin.col_begin : grid.x;
        const uint32_t c1 = spatial_bin.col_end < grid.x ? spatial_bin.col_end : grid.x;
        const uint32_t width = c1 > c0 ? c1 - c0 : 0u;
        const uint64_t cells = (uint64_t)width * grid.y;
        if (update.indices.size() != cells + 2) return false;
        const uint64_t n = update.indices[cells + 1];
        if (update.positions.size() < n || update.velocities.size() < n) return false;
        for (uint64_t i = 0; i < cells; i++) {
            const SpatialBinCoord coord{bl.x + (int32_t)(c0 + i % width), bl.y + (int32_t)(i / width)};
            const uint32_t b = update.indices[i + 1], e = update.indices[i + 2];
            if (e < b || e > n) return false;
            if (e == b) {
                hashmap.erase(coord);
                continue;
            }
            ParticleData &d = hashmap[coord];
            d.positions.assign(update.positions.begin() + b, update.positions.begin() + e);
            d.velocities.assign(update.velocities.begin() + b, update.velocities.begin() + e);
        }
        return true;
    }
    PackedData create_packed_data() {  // :90-103
        PackedData data = spatial_bin.create_packed_data(*this);
        particles_in_frame_count = (uint32_t)data.positions.size();
        return data;
    }
    uint32_t max_particles_per_frame() const {  // :116-133
        SpatialBinCoord bl;
        UVec2 grid;
        spatial_bin.get_active_cells(bl, grid);
        return wrach_host_max_particles_per_frame(grid.x * grid.y, spatial_bin.cell_size);
    }
    const std::vector<Particle> &log() const { return log_; }
    bool bucketed_empty() const { return hashmap.empty(); }
    void flush_log() {
        for (const Particle &p : log_) {
            ParticleData &e = hashmap[spatial_bin.get_cell_coord(p.position)];
            e.positions.push_back(p.position);
            e.velocities.push_back(p.velocity);
        }
        log_.clear();
    }

  private:
    std::vector<Particle> log_;  // particles added and not yet bucketed, in insertion order
};

// Page-locking of the vectors `tick` fills (see wrach_host.cpp); released with the state.
struct PinnedPackedData {
    const void *reg_ptr[3] = {nullptr, nullptr, nullptr};
    size_t reg_bytes[3] = {0, 0, 0};
    void follow(PackedData &d);
    void release();
    PinnedPackedData() = default;
    PinnedPackedData(const PinnedPackedData &) = delete;
    PinnedPackedData &operator=(const PinnedPackedData &) = delete;
    ~PinnedPackedData() { release(); }
};

struct GPUUploadSettings { ShaderWorldSettings settings; };
using GPUUpload = std::variant<PackedData, GPUUploadSettings>;  // state.rs:54-59

class WrachState {  // state.rs:17-101
  public:
    WrachConfig config;
    ShaderWorldSettings shader_settings{};
    ParticleStore particle_store;
    PackedData packed_data;
    PinnedPackedData pinned;  // declared after packed_data: released before the vectors are freed
    std::vector<GPUUpload> gpu_uploads;

    explicit WrachState(const WrachConfig &c)  // :65-80
        : config(c), particle_store(c.cell_size, Vec4{0.0f, 0.0f, (float)c.dimensions[0], (float)c.dimensions[1]}) {
        // PhysicsComputeWorker::build fills these in (compute/builder.rs:56-66)
        shader_settings.view_dimensions[0] = (float)c.dimensions[0];
        shader_settings.view_dimensions[1] = (float)c.dimensions[1];
        shader_settings.view_anchor[0] = shader_settings.view_anchor[1] = 0.0f;
        shader_settings.grid_dimensions[0] = particle_store.spatial_bin.grid_dimensions.x;
        shader_settings.grid_dimensions[1] = particle_store.spatial_bin.grid_dimensions.y;
        shader_settings.cell_size = c.cell_size;
        shader_settings.particles_in_frame_count = 0;
    }
    void gpu_upload(GPUUpload upload) { gpu_uploads.push_back(std::move(upload)); }  // :84-86
    void add_particles(const std::vector<Particle> &particles) {                     // :90-101
        for (const Particle &p : particles) particle_store.add_particle(p);
        gpu_upload(particle_store.create_packed_data());
        shader_settings.particles_in_frame_count = particle_store.particles_in_frame_count;
        gpu_upload(GPUUploadSettings{shader_settings});
    }
    // The reference's intended "window onto a larger world" (particle_store.rs:22-26, builder.rs:61
    // hard-wires the anchor to 0): write what the GPU holds back into the store, then move the
    // viewport and queue the new frame.  The worker's buffers were sized for the old grid, so the
    // new viewport must cover the same number of cell columns and rows; and because the store keys
    // particles by absolute cell (spatial_bin.rs:48-64) while the shaders key them relative to the
    // anchor (particles_per_cell.wgsl:14-15), the anchor must lie on a cell boundary.
    int update_from_gpu() { return particle_store.update_from_gpu(packed_data) ? WRACH_OK : WRACH_ERR_STATE; }
    int set_viewport(Vec4 viewport) {
        const float cs = (float)config.cell_size;
        if (!(viewport.z >= viewport.x) || !(viewport.w >= viewport.y) || std::fmod(viewport.x, cs) != 0.0f ||
            std::fmod(viewport.y, cs) != 0.0f)
            return WRACH_ERR_BAD_ARG;
        SpatialBin moved = particle_store.spatial_bin;
        moved.viewport = viewport;
        moved.update_grid_size();
        if (moved.grid_dimensions.x != particle_store.spatial_bin.grid_dimensions.x ||
            moved.grid_dimensions.y != particle_store.spatial_bin.grid_dimensions.y)
            return WRACH_ERR_BAD_ARG;
        particle_store.spatial_bin = moved;
        shader_settings.view_anchor[0] = viewport.x;
        shader_settings.view_anchor[1] = viewport.y;
        shader_settings.view_dimensions[0] = viewport.z - viewport.x;
        shader_settings.view_dimensions[1] = viewport.w - viewport.y;
        gpu_upload(particle_store.create_packed_data());
        shader_settings.particles_in_frame_count = particle_store.particles_in_frame_count;
        gpu_upload(GPUUploadSettings{shader_settings});
        return WRACH_OK;
    }
    uint32_t total_cells() const {  // compute/builder.rs:30-37
        return particle_store.spatial_bin.grid_dimensions.x * particle_store.spatial_bin.grid_dimensions.y + 2u;
    }
};

}  // namespace wrach::host
