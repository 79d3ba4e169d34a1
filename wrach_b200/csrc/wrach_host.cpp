// wrach_host.cpp — host-side systems of the Wrach plugin restated in C++ against the CUDA worker's
// C ABI, plus the C facade of include/wrach_host.h.  File:line map in that header.
#include "wrach_host.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <memory>
#include <thread>
#include <new>
#include <string>

namespace wrach::host {

// SpatialBin::create_packed_data — spatial_bin.rs:103-149.  Walks the active cells row-major,
// emits two leading zeros (first-slot marker + the prefix-sum shift, :111-120) then running totals,
// and concatenates each cell's particles in insertion order.  Cells outside the viewport stay in
// the store and are skipped (particle_store.rs:214-228).
constexpr size_t kPinnedFrom = 32u << 20;  // bytes: smaller arrays are not worth a cudaMallocHost call
void *host_array_alloc(size_t bytes) {
    if (bytes >= kPinnedFrom) {
        void *p = nullptr;
        if (cudaMallocHost(&p, bytes) == cudaSuccess) return p;
        cudaGetLastError();  // no device / no room: ordinary memory (wrach_cuda_read copies into anything)
    }
    void *p = ::operator new(bytes ? bytes : 1);
    return p;
}
void host_array_free(void *p, size_t bytes) {
    if (!p) return;
    if (bytes >= kPinnedFrom) {
        // which of the two it was: ask the runtime (cheap next to the transfer such an array is for)
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
            cudaFreeHost(p);
            return;
        }
        cudaGetLastError();
    }
    ::operator delete(p);
}

PackedData SpatialBin::create_packed_data(const ParticleStore &store) const {
    SpatialBinCoord bl;
    UVec2 grid;
    get_active_cells(bl, grid);
    const uint32_t c0 = std::min(col_begin, grid.x), c1 = std::min(col_end, grid.x);  // strip window
    const uint32_t width = c1 > c0 ? c1 - c0 : 0u;
    const uint64_t cells = (uint64_t)width * grid.y;
    auto active_index = [&](SpatialBinCoord c) -> int64_t {
        const int64_t cx = (int64_t)c.x - bl.x, cy = (int64_t)c.y - bl.y;
        if (cx < (int64_t)c0 || cy < 0 || cx >= (int64_t)c1 || cy >= (int64_t)grid.y) return -1;
        return cy * (int64_t)width + (cx - c0);
    };
    std::vector<uint32_t> cursor(cells + 1, 0u);
    for (const auto &kv : store.hashmap) {
        const int64_t a = active_index(kv.first);
        if (a >= 0) cursor[a] += (uint32_t)kv.second.positions.size();
    }
    // The log (bulk inserts not yet bucketed) is counting-sorted by cell on all host threads: thread
    // t owns a contiguous range of cells, reads the whole key array in log order (sequential) and
    // counts / scatters only the particles of its own cells, so insertion order inside a cell -- what
    // the reference's per-cell Vec gives -- is kept and no two threads touch the same counter.
    const std::vector<Particle> &log = store.log();
    const size_t n_log = log.size();
    std::vector<uint32_t> log_cell(n_log);  // 0xFFFFFFFF: outside the packed window (total_cells is a u32: cells < 2^32 - 2)
    const unsigned nt = (unsigned)std::max<uint64_t>(
        1, std::min<uint64_t>({(uint64_t)std::max(1u, std::thread::hardware_concurrency()), 64ull, n_log / 65536 + 1, cells}));
    auto parallel = [&](auto &&fn) {
        if (nt == 1) { fn(0u); return; }
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++) pool.emplace_back(fn, t);
        for (auto &th : pool) th.join();
    };
    parallel([&](unsigned t) {
        for (size_t i = n_log * t / nt, e = n_log * (t + 1) / nt; i < e; i++)
            log_cell[i] = (uint32_t)active_index(get_cell_coord(log[i].position));  // -1 -> 0xFFFFFFFF
    });
    auto cell_range = [&](unsigned t, uint32_t &lo, uint32_t &hi) {
        lo = (uint32_t)(cells * t / nt);
        hi = (uint32_t)(cells * (t + 1) / nt);
    };
    parallel([&](unsigned t) {
        uint32_t lo, hi;
        cell_range(t, lo, hi);
        for (size_t i = 0; i < n_log; i++) {
            const uint32_t c = log_cell[i];
            if (c >= lo && c < hi) cursor[c]++;
        }
    });
    PackedData out;
    out.indices.resize(cells + 2);
    out.indices[0] = 0;
    out.indices[1] = 0;
    uint32_t running = 0;
    for (uint64_t c = 0; c < cells; c++) {
        const uint32_t n = cursor[c];
        cursor[c] = running;
        running += n;
        out.indices[c + 2] = running;
    }
    out.positions.resize(running);
    out.velocities.resize(running);
    for (const auto &kv : store.hashmap) {  // bucketed particles are older than anything in the log
        const int64_t a = active_index(kv.first);
        if (a < 0) continue;
        const size_t n = kv.second.positions.size();
        std::copy_n(kv.second.positions.begin(), n, out.positions.begin() + cursor[a]);
        std::copy_n(kv.second.velocities.begin(), n, out.velocities.begin() + cursor[a]);
        cursor[a] += (uint32_t)n;
    }
    parallel([&](unsigned t) {
        uint32_t lo, hi;
        cell_range(t, lo, hi);
        for (size_t i = 0; i < n_log; i++) {
            const uint32_t c = log_cell[i];
            if (c < lo || c >= hi) continue;
            const uint32_t d = cursor[c]++;
            out.positions[d] = log[i].position;
            out.velocities[d] = log[i].velocity;
        }
    });
    return out;
}

// maybe_upload_to_gpu — plugin/build.rs:88-126
int maybe_upload_to_gpu(wrach_cuda_worker *worker, WrachState &state) {
    if (state.gpu_uploads.empty()) return WRACH_OK;
    for (const GPUUpload &upload : state.gpu_uploads) {
        int rc = WRACH_OK;
        if (const PackedData *data = std::get_if<PackedData>(&upload)) {
            if (!data->indices.empty())
                rc = wrach_cuda_write_slice(worker, WRACH_INDICES_MAIN, data->indices.data(),
                                            data->indices.size() * sizeof(uint32_t));
            if (!rc && !data->positions.empty())
                rc = wrach_cuda_write_slice(worker, WRACH_POSITIONS_IN, data->positions.data(),
                                            data->positions.size() * sizeof(Vec2));
            if (!rc && !data->velocities.empty())
                rc = wrach_cuda_write_slice(worker, WRACH_VELOCITIES_IN, data->velocities.data(),
                                            data->velocities.size() * sizeof(Vec2));
        } else {
            rc = wrach_cuda_write_settings(worker, &std::get<GPUUploadSettings>(upload).settings);
        }
        if (rc) return rc;
    }
    state.gpu_uploads.clear();
    return WRACH_OK;
}

// Page-lock the three vectors `tick` reads into, so the read-backs run at the full PCIe rate.  The
// registration follows the vectors' storage (re-done when a resize moved or grew it).
void PinnedPackedData::follow(PackedData &d) {
    const void *ptr[3] = {d.indices.data(), d.positions.data(), d.velocities.data()};
    const size_t bytes[3] = {d.indices.capacity() * sizeof(uint32_t), d.positions.capacity() * sizeof(Vec2),
                             d.velocities.capacity() * sizeof(Vec2)};
    for (int i = 0; i < 3; i++) {
        if (ptr[i] == reg_ptr[i] && bytes[i] == reg_bytes[i]) continue;
        if (reg_ptr[i]) wrach_cuda_host_unregister(const_cast<void *>(reg_ptr[i]));
        reg_ptr[i] = nullptr;
        reg_bytes[i] = 0;
        if (bytes[i] >= kPinnedFrom) {  // HostArray storage of this size is cudaMallocHost memory already, when a device exists
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, ptr[i]) == cudaSuccess && attr.type == cudaMemoryTypeHost) continue;
            cudaGetLastError();
        }
        if (ptr[i] && bytes[i] >= (1u << 16) && wrach_cuda_host_register(const_cast<void *>(ptr[i]), bytes[i]) == WRACH_OK) {
            reg_ptr[i] = ptr[i];
            reg_bytes[i] = bytes[i];
        }
    }
}
void PinnedPackedData::release() {
    for (int i = 0; i < 3; i++) {
        if (reg_ptr[i]) wrach_cuda_host_unregister(const_cast<void *>(reg_ptr[i]));
        reg_ptr[i] = nullptr;
        reg_bytes[i] = 0;
    }
}

// tick — plugin/build.rs:135-158: `if !compute_worker.ready() { return; }`, then read the three
// CPU-visible buffers back into packed_data.  Returns WRACH_OK when the frame was read,
// WRACH_TICK_SKIPPED (1) when the worker was still busy and the frame was skipped as in the
// reference, a negative status on error.  The three read_vec calls are queued copies followed by
// ONE synchronisation.
static int read_back(wrach_cuda_worker *worker, WrachState &state, size_t n_slots) {
    const size_t ib = wrach_cuda_buffer_bytes(worker, WRACH_INDICES_MAIN);
    if (state.packed_data.positions.capacity() < n_slots || state.packed_data.velocities.capacity() < n_slots)
        state.pinned.release();  // the resize below moves the storage: never free page-locked memory
    state.packed_data.positions.resize(n_slots);
    state.packed_data.velocities.resize(n_slots);
    state.pinned.follow(state.packed_data);
    int rc = WRACH_OK;
    if (n_slots == wrach_cuda_buffer_bytes(worker, WRACH_POSITIONS_IN) / sizeof(Vec2))  // tick: indices travel with the rest
        rc = wrach_cuda_read_async(worker, WRACH_INDICES_MAIN, state.packed_data.indices.data(), ib);
    if (!rc) rc = wrach_cuda_read_async(worker, WRACH_POSITIONS_IN, state.packed_data.positions.data(), n_slots * sizeof(Vec2));
    if (!rc) rc = wrach_cuda_read_async(worker, WRACH_VELOCITIES_IN, state.packed_data.velocities.data(), n_slots * sizeof(Vec2));
    if (!rc) rc = wrach_cuda_sync(worker);
    return rc;
}

int tick(wrach_cuda_worker *worker, WrachState &state, bool wait) {
    if (!wait) {
        const int ready = wrach_cuda_ready(worker);
        if (ready < 0) return ready;
        if (ready == 0) return WRACH_TICK_SKIPPED;
    }
    const size_t ib = wrach_cuda_buffer_bytes(worker, WRACH_INDICES_MAIN);
    const size_t pb = wrach_cuda_buffer_bytes(worker, WRACH_POSITIONS_IN);
    if (state.packed_data.indices.capacity() < ib / sizeof(uint32_t)) state.pinned.release();
    state.packed_data.indices.resize(ib / sizeof(uint32_t));
    return read_back(worker, state, pb / sizeof(Vec2));
}

}  // namespace wrach::host

struct wrach_state {  // the C handle of a WrachState
    wrach::host::WrachState st;
    explicit wrach_state(const wrach::host::WrachConfig &c) : st(c) {}
};

namespace wrach::host {

// tick, active part only (not in the reference): the reference reads all three
// buffers back at full capacity every frame (P slots) although only the first N hold particles.
// This variant reads `indices` first, takes N = indices[last] and fetches N slots of each array;
// packed_data.positions / velocities then have length N.
int tick_active(wrach_cuda_worker *worker, WrachState &state) {
    const size_t ib = wrach_cuda_buffer_bytes(worker, WRACH_INDICES_MAIN);
    if (state.packed_data.indices.capacity() < ib / sizeof(uint32_t)) state.pinned.release();
    state.packed_data.indices.resize(ib / sizeof(uint32_t));
    state.pinned.follow(state.packed_data);
    int rc = wrach_cuda_read(worker, WRACH_INDICES_MAIN, state.packed_data.indices.data(), ib);
    if (rc) return rc;
    const size_t n = state.packed_data.indices.empty() ? 0 : state.packed_data.indices.back();
    return read_back(worker, state, n);
}

// WrachAPI — runners/api/src/lib.rs:17-87: the plugin wired to one worker, no windowing.
class WrachAPI {
  public:
    std::unique_ptr<wrach_state> box;  // owns the WrachState resource
    WrachState &state;
    wrach_cuda_worker *worker = nullptr;
    std::vector<Vec2> positions, velocities;  // lib.rs:21-24

    explicit WrachAPI(const WrachConfig &c) : box(new wrach_state(c)), state(box->st) {}
    ~WrachAPI() { wrach_cuda_destroy(worker); }

    int init(int device, int arith) {  // WrachPlugin::build -> PhysicsComputeWorker::build
        const uint32_t total = state.total_cells();
        if (total >= (1ull << 32) - 2) return WRACH_ERR_BAD_ARG;
        return wrach_cuda_create(&state.shader_settings, total, state.particle_store.max_particles_per_frame(),
                                 device, arith, &worker);
    }
    int tick_frame() {  // lib.rs:49-52: app.update() then read_data()
        int rc = maybe_upload_to_gpu(worker, state);
        if (!rc) rc = wrach_cuda_step(worker, 1);
        if (!rc) rc = tick(worker, state, true);  // the headless API reads every frame (lib.rs:49-52)
        if (!rc) read_data();
        return rc;
    }
    void read_data() {  // lib.rs:58-75
        positions.assign(state.packed_data.positions.begin(), state.packed_data.positions.end());
        velocities.assign(state.packed_data.velocities.begin(), state.packed_data.velocities.end());
    }
};

}  // namespace wrach::host

using namespace wrach::host;

struct wrach_api {
    WrachAPI api;
    explicit wrach_api(const WrachConfig &c) : api(c) {}
};

namespace {
WrachConfig to_cpp(const wrach_config *c) {
    WrachConfig o;
    if (c) {
        o.dimensions[0] = c->dimensions[0];
        o.dimensions[1] = c->dimensions[1];
        o.cell_size = c->cell_size;
        o.boundaries_as_dimensions = c->boundaries_as_dimensions != 0;
    }
    return o;
}
int add_particles_impl(WrachState &st, const float *p, uint64_t n) {
    if (!p && n) return WRACH_ERR_BAD_ARG;
    static_assert(sizeof(Particle) == 16, "Particle is (x, y, vx, vy)");
    std::vector<Particle> v(n);
    if (n) memcpy(static_cast<void *>(v.data()), p, n * sizeof(Particle));
    st.add_particles(v);
    return WRACH_OK;
}
}  // namespace

// Self-checks of a packed frame (include/wrach_host.h): cells are dealt to the host threads in
// contiguous ranges, each thread walks the slots of its cells.
namespace {
template <typename Fn>
void for_cell_ranges(uint64_t cells, Fn &&fn) {
    const unsigned nt = (unsigned)std::max<uint64_t>(
        1, std::min<uint64_t>({(uint64_t)std::max(1u, std::thread::hardware_concurrency()), 64ull, cells / 4096 + 1}));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++) pool.emplace_back(fn, t, cells * t / nt, cells * (t + 1) / nt);
    for (auto &th : pool) th.join();
}
}  // namespace

extern "C" {

void wrach_config_default(wrach_config *out) {
    if (!out) return;
    WrachConfig d;
    out->dimensions[0] = d.dimensions[0];
    out->dimensions[1] = d.dimensions[1];
    out->cell_size = d.cell_size;
    out->boundaries_as_dimensions = 0;
    out->reserved = 0;
}

// Seeded scene generator (wrach_b200/scene.py documents the formula; that numpy version is the
// definition, this is the same arithmetic on all host threads).
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline float unit24(uint64_t seed, uint64_t id, uint64_t comp) {
    return (float)(splitmix64(seed ^ splitmix64(id * 4ull + comp)) >> 40) * (1.0f / 16777216.0f);
}
void wrach_host_generate_scene(uint64_t seed, uint64_t first_id, uint64_t n, float x0, float width, float height,
                               int pile, float *out) {
    unsigned nt = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++)
        pool.emplace_back([=] {
            for (uint64_t i = n * t / nt, e = n * (t + 1) / nt; i < e; i++) {
                const uint64_t id = first_id + i;
                float uy = unit24(seed, id, 1);
                if (pile) uy = (uy * uy) * (uy * uy);
                out[4 * i + 0] = x0 + unit24(seed, id, 0) * width;
                out[4 * i + 1] = uy * height;
                out[4 * i + 2] = unit24(seed, id, 2) - 0.5f;
                out[4 * i + 3] = unit24(seed, id, 3) - 0.5f;
            }
        });
    for (auto &th : pool) th.join();
}

int wrach_host_check_packed(const uint32_t *ind, uint64_t n_ind, const float *pos, const float *vel, uint32_t c0,
                            uint32_t c1, uint32_t grid_x, float width_f, float height_f, uint16_t cell_size) {
    if (!ind || n_ind < 2 || c1 <= c0 || c1 > grid_x || ((!pos || !vel) && ind[n_ind - 1])) return WRACH_ERR_BAD_ARG - 100;
    const uint64_t cells = n_ind - 2, width = c1 - c0;
    if (ind[0] != 0 || cells % width) return -1;
    std::vector<int> worst(65, 0);
    for_cell_ranges(cells, [&](unsigned t, uint64_t lo, uint64_t hi) {
        int bad = 0;
        const float cs = (float)cell_size;
        for (uint64_t c = lo; c < hi && !bad; c++) {
            const uint32_t b = ind[c + 1], e = ind[c + 2];
            if (e < b) { bad = -1; break; }
            const int64_t cx = (int64_t)(c0 + c % width), cy = (int64_t)(c / width);
            for (uint32_t j = b; j < e; j++) {
                const float x = pos[2 * (size_t)j], y = pos[2 * (size_t)j + 1];
                if (!(x >= 0.0f && x <= width_f && y >= 0.0f && y <= height_f)) { bad = -2; break; }
                if (!(std::fabs(vel[2 * (size_t)j]) <= 1.0f && std::fabs(vel[2 * (size_t)j + 1]) <= 1.0f)) { bad = -3; break; }
                const int64_t px = div_euclid_as_i32(x, cs), py = div_euclid_as_i32(y, cs);
                if (px < (int64_t)c0 || px >= (int64_t)c1) { bad = -4; break; }
                if (px != cx || py != cy) { bad = -5; break; }
            }
        }
        worst[t] = bad;
    });
    if (cells && ind[1] != 0) return -1;
    for (int b : worst)
        if (b) return b;
    return 0;
}

uint64_t wrach_host_packed_checksum(const uint32_t *ind, uint64_t n_ind, const float *pos, const float *vel, uint32_t c0,
                                    uint32_t c1, uint32_t grid_x) {
    if (!ind || n_ind < 2 || c1 <= c0) return 0;
    const uint64_t cells = n_ind - 2, width = c1 - c0;
    std::vector<uint64_t> part(65, 0);
    const uint32_t *pb = reinterpret_cast<const uint32_t *>(pos), *vb = reinterpret_cast<const uint32_t *>(vel);
    for_cell_ranges(cells, [&](unsigned t, uint64_t lo, uint64_t hi) {
        uint64_t sum = 0;
        for (uint64_t c = lo; c < hi; c++) {
            const uint64_t gcell = (c / width) * grid_x + c0 + c % width;
            const uint32_t b = ind[c + 1], e = ind[c + 2];
            for (uint32_t j = b; j < e; j++) {
                uint64_t h = splitmix64(splitmix64(gcell) ^ (uint64_t)(j - b));
                h = splitmix64(h ^ ((uint64_t)pb[2 * (size_t)j] << 32 | pb[2 * (size_t)j + 1]));
                h = splitmix64(h ^ ((uint64_t)vb[2 * (size_t)j] << 32 | vb[2 * (size_t)j + 1]));
                sum += h;
            }
        }
        part[t] = sum;
    });
    uint64_t total = 0;
    for (uint64_t v : part) total += v;
    return total;
}

int32_t wrach_host_cell_coord(float position, uint16_t cell_size) {
    return div_euclid_as_i32(position, (float)cell_size);
}

void wrach_host_active_grid(const float viewport[4], uint16_t cell_size, int32_t bottom_left[2], uint32_t grid[2]) {
    SpatialBin bin(cell_size, Vec4{viewport[0], viewport[1], viewport[2], viewport[3]});
    SpatialBinCoord bl;
    UVec2 g;
    bin.get_active_cells(bl, g);
    bottom_left[0] = bl.x; bottom_left[1] = bl.y;
    grid[0] = g.x; grid[1] = g.y;
}

uint32_t wrach_host_max_particles_per_frame(uint32_t total_cells, uint16_t cell_size) {
    const uint32_t per_cell = (uint32_t)cell_size * (uint32_t)cell_size;  // particle_store.rs:127
    const uint32_t normally = total_cells * per_cell;
    const uint32_t one_percent = (normally + 99u) / 100u;                 // div_ceil(100)
    return normally + 10u * one_percent;                                  // extra_percent = 10
}

wrach_state *wrach_state_new(const wrach_config *config) { return new (std::nothrow) wrach_state(to_cpp(config)); }
wrach_state *wrach_state_new_strip(const wrach_config *config, uint32_t col_begin, uint32_t col_end) {
    wrach_state *s = wrach_state_new(config);
    if (s) {
        s->st.particle_store.spatial_bin.col_begin = col_begin;
        s->st.particle_store.spatial_bin.col_end = col_end;
    }
    return s;
}
void wrach_state_free(wrach_state *s) { delete s; }

int wrach_state_add_particles(wrach_state *s, const float *p, uint64_t n) {
    return s ? add_particles_impl(s->st, p, n) : WRACH_ERR_BAD_ARG;
}
uint32_t wrach_state_pending_uploads(const wrach_state *s) { return s ? (uint32_t)s->st.gpu_uploads.size() : 0u; }
void wrach_state_shader_settings(const wrach_state *s, wrach_world_settings *out) {
    if (s && out) *out = s->st.shader_settings;
}
void wrach_state_grid(const wrach_state *s, uint32_t grid[2], uint32_t *total_cells, uint32_t *max_particles) {
    if (!s) return;
    if (grid) {
        grid[0] = s->st.particle_store.spatial_bin.grid_dimensions.x;
        grid[1] = s->st.particle_store.spatial_bin.grid_dimensions.y;
    }
    const SpatialBin &bin = s->st.particle_store.spatial_bin;
    const uint32_t width = std::min(bin.col_end, bin.grid_dimensions.x) - std::min(bin.col_begin, bin.grid_dimensions.x);
    if (total_cells) *total_cells = width * bin.grid_dimensions.y + 2u;  // == total_cells() without a strip window
    if (max_particles) *max_particles = wrach_host_max_particles_per_frame(width * bin.grid_dimensions.y, bin.cell_size);
}
const uint32_t *wrach_state_packed_indices(const wrach_state *s, uint64_t *len) {
    if (len) *len = s ? s->st.packed_data.indices.size() : 0;
    return s ? s->st.packed_data.indices.data() : nullptr;
}
const float *wrach_state_packed_positions(const wrach_state *s, uint64_t *len) {
    if (len) *len = s ? s->st.packed_data.positions.size() : 0;
    return s ? reinterpret_cast<const float *>(s->st.packed_data.positions.data()) : nullptr;
}
const float *wrach_state_packed_velocities(const wrach_state *s, uint64_t *len) {
    if (len) *len = s ? s->st.packed_data.velocities.size() : 0;
    return s ? reinterpret_cast<const float *>(s->st.packed_data.velocities.data()) : nullptr;
}
uint32_t wrach_state_create_packed_data(wrach_state *s, uint32_t *indices, float *positions, float *velocities) {
    if (!s) return 0;
    PackedData d = s->st.particle_store.create_packed_data();
    if (indices) memcpy(indices, d.indices.data(), d.indices.size() * sizeof(uint32_t));
    if (positions) memcpy(positions, d.positions.data(), d.positions.size() * sizeof(Vec2));
    if (velocities) memcpy(velocities, d.velocities.data(), d.velocities.size() * sizeof(Vec2));
    return (uint32_t)d.positions.size();
}

int wrach_state_set_packed_data(wrach_state *s, const uint32_t *indices, uint64_t n_indices, const float *positions,
                                const float *velocities, uint64_t n) {
    if (!s || (!indices && n_indices) || ((!positions || !velocities) && n)) return WRACH_ERR_BAD_ARG;
    PackedData &d = s->st.packed_data;
    s->st.pinned.release();  // the assignments below may move the storage
    d.indices.assign(indices, indices + n_indices);
    d.positions.resize(n);
    d.velocities.resize(n);
    if (n) {
        memcpy(static_cast<void *>(d.positions.data()), positions, n * sizeof(Vec2));
        memcpy(static_cast<void *>(d.velocities.data()), velocities, n * sizeof(Vec2));
    }
    return WRACH_OK;
}
int wrach_host_strip_exchange_plan(uint32_t n_ranks, uint32_t rank, const uint32_t *rows, uint32_t stride,
                                   uint32_t *send_off, uint32_t *recv_off, uint32_t *n_recv) {
    int over = -1;
    for (uint32_t d = 0; d < n_ranks && over < 0; d++) {
        uint64_t arriving = 0;
        for (uint32_t s_ = 0; s_ < n_ranks; s_++) arriving += rows[(size_t)s_ * stride + d];
        if (arriving > rows[(size_t)d * stride + n_ranks]) over = (int)d;
    }
    uint32_t off = 0;
    for (uint32_t d = 0; d < n_ranks; d++) {
        if (send_off) send_off[d] = off;
        off += rows[(size_t)rank * stride + d];
    }
    off = 0;
    for (uint32_t s_ = 0; s_ < n_ranks; s_++) {
        if (recv_off) recv_off[s_] = off;
        off += rows[(size_t)s_ * stride + rank];
    }
    if (n_recv) *n_recv = off;
    return over;
}
int wrach_state_update_from_gpu(wrach_state *s) { return s ? s->st.update_from_gpu() : WRACH_ERR_BAD_ARG; }
int wrach_state_set_viewport(wrach_state *s, const float viewport[4]) {
    return (s && viewport) ? s->st.set_viewport(Vec4{viewport[0], viewport[1], viewport[2], viewport[3]})
                           : WRACH_ERR_BAD_ARG;
}
uint64_t wrach_state_stored_particles(const wrach_state *s) {
    if (!s) return 0;
    uint64_t n = s->st.particle_store.log().size();
    for (const auto &kv : s->st.particle_store.hashmap) n += kv.second.positions.size();
    return n;
}

int wrach_plugin_maybe_upload_to_gpu(wrach_cuda_worker *worker, wrach_state *s) {
    return (worker && s) ? maybe_upload_to_gpu(worker, s->st) : WRACH_ERR_BAD_ARG;
}
int wrach_plugin_tick_active(wrach_cuda_worker *worker, wrach_state *s) {
    return (worker && s) ? tick_active(worker, s->st) : WRACH_ERR_BAD_ARG;
}
int wrach_plugin_tick(wrach_cuda_worker *worker, wrach_state *s) {
    return (worker && s) ? tick(worker, s->st, false) : WRACH_ERR_BAD_ARG;
}
int wrach_plugin_tick_wait(wrach_cuda_worker *worker, wrach_state *s) {
    return (worker && s) ? tick(worker, s->st, true) : WRACH_ERR_BAD_ARG;
}

int wrach_api_new(const wrach_config *config, int device, int arith, wrach_api **out) {
    if (!out) return WRACH_ERR_BAD_ARG;
    *out = nullptr;
    wrach_api *a = new (std::nothrow) wrach_api(to_cpp(config));
    if (!a) return WRACH_ERR_BAD_ARG;
    int rc = a->api.init(device, arith);
    if (rc) {
        delete a;
        return rc;
    }
    *out = a;
    return WRACH_OK;
}
void wrach_api_free(wrach_api *a) { delete a; }
int wrach_api_tick(wrach_api *a) { return a ? a->api.tick_frame() : WRACH_ERR_BAD_ARG; }
int wrach_api_add_particles(wrach_api *a, const float *p, uint64_t n) {
    return a ? add_particles_impl(a->api.state, p, n) : WRACH_ERR_BAD_ARG;
}
const float *wrach_api_positions(const wrach_api *a, uint64_t *len) {
    if (len) *len = a ? a->api.positions.size() : 0;
    return a ? reinterpret_cast<const float *>(a->api.positions.data()) : nullptr;
}
const float *wrach_api_velocities(const wrach_api *a, uint64_t *len) {
    if (len) *len = a ? a->api.velocities.size() : 0;
    return a ? reinterpret_cast<const float *>(a->api.velocities.data()) : nullptr;
}
wrach_state *wrach_api_get_simulation_state(wrach_api *a) {
    return a ? a->api.box.get() : nullptr;  // owned by the API object
}
wrach_cuda_worker *wrach_api_worker(wrach_api *a) { return a ? a->api.worker : nullptr; }
const char *wrach_api_last_error(const wrach_api *a) {
    return a && a->api.worker ? wrach_cuda_last_error(a->api.worker) : wrach_cuda_last_error(nullptr);
}

}  // extern "C"
