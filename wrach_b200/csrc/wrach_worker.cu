// wrach_worker.cu — the C ABI of include/wrach_cuda.h: buffer set, uploads, step sequencing,
// read-back.  Host-side twin of PhysicsComputeWorker::build (runners/bevy/src/compute/builder.rs:24-92)
// and of the bevy_easy_compute worker calls Wrach makes (runners/bevy/src/plugin/build.rs:88-158).
//
// No CPU fallback lives here: every compute call is a CUDA kernel from wrach_kernels.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "wrach_kernels.cuh"
#include "wrach_tiles.cuh"
#include "wrach_xrebin.cuh"
#include "../../include/wrach_host.h"

static_assert(sizeof(wrach_world_settings) == 32, "uniform must be 32 bytes (config_shader.rs:15-29)");
static_assert(offsetof(wrach_world_settings, view_dimensions) == 0, "layout");
static_assert(offsetof(wrach_world_settings, view_anchor) == 8, "layout");
static_assert(offsetof(wrach_world_settings, grid_dimensions) == 16, "layout");
static_assert(offsetof(wrach_world_settings, cell_size) == 24, "layout");
static_assert(offsetof(wrach_world_settings, particles_in_frame_count) == 28, "layout");

using namespace wrach;

struct wrach_cuda_worker {
    std::mutex mu;
    int device = 0;
    bool pdl = WRACH_PDL != 0;      // programmatic dependent launch between the frame's kernels (WRACH_PDL=0/1 in the
                                    // environment overrides the build's default: A/B runs on one box)
    bool neighbour_mode = false;    // opt-in 3x3 neighbour search before k_phys (an extension: the reference has none)
    bool pdl_forced = false, pdl_active = false;
    uint32_t resident_phys_blocks = 148 * WRACH_PHYS_MINBLOCKS;  // blocks of k_phys the device holds at once
    bool dense_enabled = false;     // a frame has taken the general path: k_rebin_dense is part of every frame
    uint32_t dense_grid = 148 * kDenseBlocksPerSM;  // blocks of k_rebin_dense: all resident
    int arith = WRACH_ARITH_SPV;
    wrach_world_settings s{};
    uint32_t total_cells = 0, cells = 0, capacity = 0;
    cudaStream_t stream = nullptr;
    uint32_t *idx[2] = {nullptr, nullptr};  // indices_main / indices_block_sums, roles swap each frame
    int cur = 0;                            // idx[cur] is INDICES_MAIN as of the last resolved frame
    float2 *pos_in = nullptr, *vel_in = nullptr, *pos_out = nullptr, *vel_out = nullptr;
    uint32_t *meta = nullptr, *cls = nullptr, *cls9 = nullptr, *goff9 = nullptr;
    uint4 *dense_list = nullptr;
    uint32_t *run_total = nullptr, *run_base = nullptr;
    uint32_t *vl_slot = nullptr;
    uint16_t *vl_meta = nullptr, *vl_cnt = nullptr;
    Ctrl *ctrl = nullptr;
    Ctrl *h_ctrl = nullptr;  // pinned mirror
    unsigned long long *tile_status = nullptr;
    uint32_t n_status = 0;
    uint32_t *slow_cursor = nullptr, *slow_src = nullptr, *slow_ticket = nullptr;  // generic re-bin scratch
    uint32_t epoch = 0;
    uint64_t pending = 0;         // frames enqueued and not yet known to have completed
    uint32_t steps_done_seen = 0; // ctrl->steps_done at the last resolve
    int cur_enqueue = 0;          // idx role the NEXT enqueued frame reads, assuming no abort
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    wrach_cuda_stats stats{};
    std::string err;
    bool dead = false;            // a fatal error left the device state unknown: no further steps
    std::string dead_why;
    // ---- fused tile frames (wrach_tiles.cuh): the state lives tile-major between read-backs
    bool tiles_on = true;            // WRACH_TILES=0 in the environment keeps every frame on k_phys / k_rebin
    bool tiles_off_until_upload = false;  // the scene does not fit the tiles (density): old path until new data arrives
    uint64_t tiles_retry_at = 0;     // a far mover sent a frame to the old path: tiles again once this many frames are done
    bool packed_valid = true;        // indices / positions_in / velocities_in hold the current state
    bool tiled_valid = false;        // tdata[tcur] / tstarts[tcur] hold the current state
    int tile_cfg = 0;
    uint32_t ntx = 0, nty = 0, ntiles = 0, tss = 0, tcap = 0;
    uint32_t t_ghost_l = 0, t_own_tc = 0, t_gx = 0;  // strips: left ghost column (0/1), owned tile columns, tile-grid width in cells
    bool strip_tiles_ok = false;     // strips: the columns were cut on tile boundaries
    bool tile_col_major = false;
    cudaStream_t comm_stream = nullptr;  // strips: the ghost exchange runs beside the interior tile columns
    cudaEvent_t ev_edge = nullptr, ev_exch = nullptr;
    uint32_t *d_sig = nullptr;       // [0]: edge blocks done (counted by the kernels), [1]: ghost exchanges delivered (written by the exchange stream)
    uint32_t edge_cum = 0, ghost_seq = 0;
    uint32_t h_count = 0;
    float4 *tdata[2] = {nullptr, nullptr};
    uint16_t *tstarts[2] = {nullptr, nullptr};
    int tcur = 0;                    // buffer the NEXT enqueued tile frame reads
    uint64_t tile_pending = 0;       // tile frames enqueued and not yet known to have completed
    uint32_t tile_ord = 0;           // ordinal of the next tile frame
    uint32_t tile_first_ord = 0;     // ordinal of the first pending one, the buffer it reads,
    int tile_first_buf = 0;
    bool tile_first_from_packed = false;  // and whether the packed state was (and stays) valid as its input
    // ---- strip workers
    bool strip = false;
    int rank = 0, n_ranks = 1;
    wrach_world_settings gs{};               // the GLOBAL world (s holds the local grid)
    uint32_t col0 = 0, col1 = 0;             // global cell columns [col0, col1) owned by this strip
    uint32_t edge_mask = 0, exp_cap = 0;
    uint8_t *exp_buf[2] = {nullptr, nullptr}, *imp_buf[2] = {nullptr, nullptr};
    uint32_t *imp_cnt = nullptr, *imp_off = nullptr;
    ncclComm_t comm = nullptr;               // NCCL mode (one process per GPU)
    wrach_cuda_worker *peer[2] = {nullptr, nullptr};  // in-process mode (wrach_cuda_strip_group_step)
    uint32_t col_end[kMaxStrips] = {};       // global column where every strip of the world ends
    // collective re-bin of one frame over all strips (wrach_xrebin.cuh); scratch lives only while it runs
    uint32_t *xr_cnt = nullptr;              // device: counts[64], cursors[64], gathered rows[64][65]
    uint32_t *xr_host = nullptr;             // pinned: the gathered rows
    XRec *xr_send = nullptr, *xr_recv = nullptr;
    unsigned long long *xr_key = nullptr;
    uint32_t *tile_vote = nullptr;           // device word: ~(ordinal + 1) of the first failed tile frame of any strip, else 0
    uint32_t *h_tile_vote = nullptr;         // pinned mirror
    // the packed buffers keep the state of the last upload / read-back while tile frames run: a checkpoint
    uint64_t ckpt_at = 0;                    // frames completed when the packed copy was last current
    // opt-in neighbour mode on strips: first-nine positions of the edge columns, sent / received per frame
    uint8_t *nb_send[2] = {nullptr, nullptr}, *nb_recv[2] = {nullptr, nullptr};
    // buffers moved into shareable allocations for a renderer (wrach_cuda_export_buffer_fd): [0] positions_in, [1] velocities_in
    struct Shared {
        CUmemGenericAllocationHandle handle = 0;
        CUdeviceptr va = 0;
        size_t bytes = 0;
    } shared[2];
};

namespace {

thread_local std::string g_create_error;

int fail(wrach_cuda_worker *w, int code, const char *fmt, ...);
// A failure that leaves frames half-enqueued or the exchange state unknown: the handle refuses
// further steps (WRACH_ERR_STATE) instead of computing on from a corrupt state; reads still work.
int die(wrach_cuda_worker *w, int code) {
    w->dead = true;
    w->dead_why = w->err;
    return code;
}

int fail(wrach_cuda_worker *w, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (w) w->err = buf; else g_create_error = buf;
    return code;
}

// NCCL is bound lazily (dlopen) so that the library loads without it and shares the copy torch
// already mapped when the caller is a torch.distributed process.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) return;
#define BIND(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.handle, "nccl" #name))
        BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(GroupStart); BIND(GroupEnd);
        BIND(Send); BIND(Recv); BIND(AllReduce); BIND(AllGather); BIND(GetErrorString);
#undef BIND
        if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllReduce || !api.AllGather || !api.GroupStart || !api.GroupEnd)
            api.handle = nullptr;
    });
    return api.handle ? &api : nullptr;
}

#define NC(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess)                                                                         \
            return fail(w, WRACH_ERR_NCCL, "%s failed: %s", #call,                                     \
                        nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "nccl error");  \
    } while (0)

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(w, WRACH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// The grid must cover every cell a clamped position can key to (particle.rs:46-70 keeps positions in
// [anchor, anchor+dims]); otherwise the reference itself would index out of bounds.
int validate_settings(wrach_cuda_worker *w, const wrach_world_settings &s, uint32_t total_cells, uint32_t capacity) {
    if (s.cell_size == 0 || s.grid_dimensions[0] == 0 || s.grid_dimensions[1] == 0)
        return fail(w, WRACH_ERR_BAD_ARG, "cell_size and grid_dimensions must be non-zero");
    const uint64_t cells = (uint64_t)s.grid_dimensions[0] * s.grid_dimensions[1];
    if (cells >= (1ull << 30))
        return fail(w, WRACH_ERR_BAD_ARG, "grid of %llu cells: at most 2^30 - 1 are supported", (unsigned long long)cells);
    if (cells + 2 != total_cells)
        return fail(w, WRACH_ERR_BAD_ARG, "total_cells (%u) != grid.x*grid.y + 2 (%llu)", total_cells,
                    (unsigned long long)cells + 2);
    const float cs = (float)s.cell_size;
    for (int a = 0; a < 2; a++) {
        if (!(s.view_dimensions[a] >= 0.0f)) return fail(w, WRACH_ERR_BAD_ARG, "view_dimensions must be >= 0");
        // the move classification compares positions relative to the anchor with exact multiples of the
        // cell size, which equals the reference's floor((x - anchor) / cell_size) below 2^23
        // (tests/test_host_mirror.py::test_fast_key_equals_divide_key_sweep); the reference's own
        // dimensions are u16 (config_app.rs:13)
        if (!(s.view_dimensions[a] < 8388608.0f))
            return fail(w, WRACH_ERR_BAD_ARG, "view_dimensions[%d]=%g: at most 2^23 - 1 is supported", a, (double)s.view_dimensions[a]);
        const float far_edge = (s.view_anchor[a] + s.view_dimensions[a]) - s.view_anchor[a];
        const float c = floorf(far_edge / cs);
        if (!(c < (float)s.grid_dimensions[a]))
            return fail(w, WRACH_ERR_BAD_ARG, "grid_dimensions[%d]=%u does not cover the viewport (needs %g)", a,
                        s.grid_dimensions[a], (double)c + 1.0);
    }
    if (s.particles_in_frame_count > capacity)
        return fail(w, WRACH_ERR_CAPACITY, "particles_in_frame_count %u > capacity %u",
                    s.particles_in_frame_count, capacity);
    return WRACH_OK;
}

Frame make_frame(wrach_cuda_worker *w, int read_role) {
    Frame f;
    f.s = w->s;
    f.lim = make_limits(w->s);
    // Programmatic dependent launch pays once a kernel is several waves of blocks long (16 M world:
    // -2 % of the frame); on a world that fits the GPU in one wave each launch edge costs more than
    // the overlap returns (1 M world: +9 %).
    // k_phys lets the next kernel's blocks in early only when that kernel is ours (k_run_scan): an
    // NCCL send/recv launched behind it must never start before the export messages are complete.
    w->pdl_active = w->pdl && (w->pdl_forced || (w->cells + kRun - 1) / kRun >= 3u * w->resident_phys_blocks);
    f.pdl = (w->pdl_active && !w->comm) ? 1u : 0u;
    f.cells = w->cells;
    f.n = w->s.particles_in_frame_count;
    f.starts = w->idx[read_role];
    f.starts_next = w->idx[read_role ^ 1];
    f.pos_in = w->pos_in;
    f.vel_in = w->vel_in;
    f.pos_out = w->pos_out;
    f.vel_out = w->vel_out;
    f.meta = w->meta;
    f.cls = w->cls;
    f.cls9 = w->cls9;
    f.goff9 = w->goff9;
    f.dense_list = w->dense_list;
    f.dense_enabled = w->dense_enabled ? 1u : 0u;
    f.run_total = w->run_total;
    f.run_base = w->run_base;
    f.vl_slot = w->vl_slot;
    f.vl_meta = w->vl_meta;
    f.vl_cnt = w->vl_cnt;
    f.ctrl = w->ctrl;
    f.tile_status = w->tile_status;
    f.epoch = ++w->epoch;
    f.col0 = w->col0;
    f.edge_mask = w->edge_mask;
    f.exp_cap = w->exp_cap;
    f.capacity = w->strip ? w->capacity : 0u;
    for (int i = 0; i < 2; i++) {
        f.exp_buf[i] = w->exp_buf[i];
        f.imp_buf[i] = w->imp_buf[i];
    }
    f.imp_cnt = w->imp_cnt;
    f.imp_off = w->imp_off;
    for (int i = 0; i < 2; i++) f.nb_halo[i] = (w->neighbour_mode && ((w->edge_mask >> i) & 1u)) ? w->nb_recv[i] : nullptr;
    return f;
}

// Launch with the programmatic-stream-serialization attribute (programmatic dependent launch): the
// kernel's blocks may become resident once every block of the kernel before it in the stream has
// called pdl_trigger (or exited), and each kernel calls pdl_wait before it touches anything that
// kernel produces (wrach_kernels.cuh).  Behind anything that never triggers -- a copy, an NCCL
// kernel, k_rebin_dense -- the attribute changes nothing.
template <typename Kernel>
void launch_frame_kernel(wrach_cuda_worker *w, Kernel kernel, uint32_t grid, uint32_t block, const Frame &f) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = w->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = w->pdl_active ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, f);  // errors surface at the cudaGetLastError that ends every enqueue
    w->stats.kernel_launches++;
}

void launch_phys(wrach_cuda_worker *w, const Frame &f) {
    const uint32_t grid = (w->cells + kRun - 1) / kRun;
    if (w->arith == WRACH_ARITH_SPV)
        launch_frame_kernel(w, k_phys<WRACH_ARITH_SPV>, grid, kRun, f);
    else
        launch_frame_kernel(w, k_phys<WRACH_ARITH_UNFUSED>, grid, kRun, f);
}

// Neighbour mode on strips: every frame starts by handing the first-nine positions of the two edge
// columns to the neighbouring strips (their k_neighbours reads them as its ghost columns).
void launch_nb_halo_pack(wrach_cuda_worker *w, const Frame &f) {
    if (!w->edge_mask) return;
    k_nb_halo_pack<<<64, 256, 0, w->stream>>>(f, (w->edge_mask & 1u) ? w->nb_send[0] : nullptr, (w->edge_mask & 2u) ? w->nb_send[1] : nullptr);
    w->stats.kernel_launches++;
}
int nb_halo_exchange_nccl(wrach_cuda_worker *w) {
    if (!w->edge_mask || !w->comm) return WRACH_OK;
    NcclApi *nc = nccl_api();
    const size_t bytes = nb_halo_bytes(w->s.grid_dimensions[1]);
    NC(nc->GroupStart());
    for (int i = 0; i < 2; i++) {
        if (!((w->edge_mask >> i) & 1u)) continue;
        const int other = w->rank + (i == 0 ? -1 : 1);
        NC(nc->Send(w->nb_send[i], bytes, ncclUint8, other, w->comm, w->stream));
        NC(nc->Recv(w->nb_recv[i], bytes, ncclUint8, other, w->comm, w->stream));
        w->stats.halo_bytes_sent += bytes;
    }
    NC(nc->GroupEnd());
    return WRACH_OK;
}
int nb_halo_exchange_peers(wrach_cuda_worker *w) {  // in-process strips: the neighbours' packs have completed
    const size_t bytes = nb_halo_bytes(w->s.grid_dimensions[1]);
    for (int i = 0; i < 2; i++) {
        if (!w->peer[i]) continue;
        CU(cudaMemcpyAsync(w->nb_recv[i], w->peer[i]->nb_send[i ^ 1], bytes, cudaMemcpyDefault, w->stream));
        w->stats.halo_bytes_sent += bytes;
    }
    return WRACH_OK;
}

// Opt-in extension (wrach_cuda_set_neighbour_mode): the cross-cell pushes, then the copy-back.
void launch_neighbours(wrach_cuda_worker *w, const Frame &f) {
    const uint64_t threads = (uint64_t)w->cells * kMaxInCell;
    const uint32_t grid_commit = (uint32_t)((threads + 255) / 256);
    const uint32_t grid = neighbour_blocks_per_row(w->s.grid_dimensions[0]) * w->s.grid_dimensions[1];
    if (grid == 0 || grid_commit == 0) return;
    if (w->arith == WRACH_ARITH_SPV)
        k_neighbours<WRACH_ARITH_SPV><<<grid, kNbThreads, 0, w->stream>>>(f);
    else
        k_neighbours<WRACH_ARITH_UNFUSED><<<grid, kNbThreads, 0, w->stream>>>(f);
    k_neighbours_commit<<<grid_commit, 256, 0, w->stream>>>(f);
    w->stats.kernel_launches += 2;
}

void launch_rebin(wrach_cuda_worker *w, const Frame &f) {
    const uint32_t grid = (w->cells + kRun - 1) / kRun;
    if (w->edge_mask) {
        k_import_index<<<32, 256, 0, w->stream>>>(f);
        w->stats.kernel_launches++;
    }
    launch_frame_kernel(w, k_run_scan, 1, 1024, f);
    launch_frame_kernel(w, k_rebin, grid, kRun, f);
    if (w->dense_enabled) {
        k_rebin_dense<<<w->dense_grid, kRun, 0, w->stream>>>(f);
        w->stats.kernel_launches++;
    }
    if (w->edge_mask) {
        k_import_place<<<32, 256, 0, w->stream>>>(f);
        w->stats.kernel_launches++;
    }
}

// strips, NCCL mode: swap the fixed-size exchange messages with both neighbours (after k_phys)
int strip_exchange_nccl(wrach_cuda_worker *w) {
    if (!w->edge_mask || !w->comm) return WRACH_OK;
    NcclApi *nc = nccl_api();
    const size_t bytes = msg_bytes(w->exp_cap);
    NC(nc->GroupStart());
    for (int i = 0; i < 2; i++) {
        if (!((w->edge_mask >> i) & 1u)) continue;
        const int other = w->rank + (i == 0 ? -1 : 1);
        NC(nc->Send(w->exp_buf[i], bytes, ncclUint8, other, w->comm, w->stream));
        NC(nc->Recv(w->imp_buf[i], bytes, ncclUint8, other, w->comm, w->stream));
        w->stats.halo_bytes_sent += bytes;
    }
    NC(nc->GroupEnd());
    // A frame one strip cannot re-bin on the fast path (a far mover, more leavers than a message
    // holds) is one NO strip re-bins: the flag is reduced over all of them before any re-bin kernel
    // reads it, so every strip stops at the same frame with its post-physics state intact and
    // resolve() runs the collective re-bin (wrach_xrebin.cuh) everywhere.
    NC(nc->AllReduce(&w->ctrl->far_seen, &w->ctrl->far_seen, 1, ncclUint32, ncclMax, w->comm, w->stream));
    return WRACH_OK;
}


// ---------------------------------------------------------------------------------------------
// Fused tile frames (wrach_tiles.cuh).  One tile shape is compiled in: 30 x 14 cells, whose ring of
// 32 x 16 staged cells is one cell per thread of a 512-thread block; a region holds 3328 slots and
// the stage 4096 particles (the 16 M benchmark scene averages 2832 / 3452: +17 % / +19 % of slack,
// nine standard deviations of a uniform scene).  Denser scenes stay on k_phys / k_rebin.
#ifndef WRACH_TILE_W
#define WRACH_TILE_W 22
#define WRACH_TILE_H 14
#define WRACH_TILE_NT 384
#define WRACH_TILE_PCAP 3072
#define WRACH_TILE_TCAP 2496
#define WRACH_TILE_MINB 3
#endif
struct TileShape {
    static constexpr int TW = WRACH_TILE_W, TH = WRACH_TILE_H, NT = WRACH_TILE_NT, PCAP = WRACH_TILE_PCAP, MINB = WRACH_TILE_MINB;
    static constexpr uint32_t TCAP = WRACH_TILE_TCAP;
};
using TileS = TileSmem<TileShape::TW, TileShape::TH, TileShape::PCAP>;
template <int ARITH, bool STRIP>
constexpr auto tile_kernel = k_tile_frame<ARITH, TileShape::TW, TileShape::TH, TileShape::NT, TileShape::PCAP, TileShape::MINB, STRIP>;

int resolve(wrach_cuda_worker *w);
int enqueue_frames(wrach_cuda_worker *w, uint64_t n, bool profile, float *phys_ms, float *rebin_ms);

bool tiles_usable(const wrach_cuda_worker *w) {
    return w->tiles_on && !w->neighbour_mode && !w->tiles_off_until_upload && w->cells > 0 &&
           (!w->strip || w->strip_tiles_ok) && w->stats.steps_completed + w->pending >= w->tiles_retry_at;
}

// Tile grid of this worker.  Single device: the grid itself.  Strip worker: its own columns (which
// wrach_cuda_strip_columns cuts on tile boundaries) plus one ghost tile column towards every
// neighbouring strip, tiles numbered column by column.
int tiles_allocate(wrach_cuda_worker *w) {
    const uint32_t gx = w->s.grid_dimensions[0], gy = w->s.grid_dimensions[1];
    const uint32_t TW = TileShape::TW;
    const uint32_t ghost_l = (w->strip && (w->edge_mask & 1u)) ? 1u : 0u, ghost_r = (w->strip && (w->edge_mask & 2u)) ? 1u : 0u;
    const uint32_t own_tc = (gx + TW - 1) / TW;
    const uint32_t ntx = ghost_l + own_tc + ghost_r, nty = (gy + TileShape::TH - 1) / TileShape::TH;
    if (w->tdata[0] && ntx == w->ntx && nty == w->nty) return WRACH_OK;
    for (int i = 0; i < 2; i++) {
        cudaFree(w->tdata[i]);
        cudaFree(w->tstarts[i]);
        w->tdata[i] = nullptr;
        w->tstarts[i] = nullptr;
    }
    w->ntx = ntx;
    w->nty = nty;
    w->ntiles = ntx * nty;
    w->t_ghost_l = ghost_l;
    w->t_own_tc = own_tc;
    w->t_gx = ghost_l * TW + gx + (ghost_r ? std::min(TW, w->gs.grid_dimensions[0] - w->col1) : 0u);
    w->tcap = TileShape::TCAP;
    w->tss = (uint32_t)((TileShape::TW * TileShape::TH + 1 + 7) & ~7);
    if ((uint64_t)w->ntiles * w->tcap >= (1ull << 32)) {  // slots are addressed with 32 bits
        w->tiles_on = false;
        return WRACH_OK;
    }
    for (int i = 0; i < 2; i++) {
        CU(cudaMalloc(&w->tdata[i], (size_t)w->ntiles * w->tcap * sizeof(float4)));
        CU(cudaMalloc(&w->tstarts[i], (size_t)w->ntiles * w->tss * sizeof(uint16_t)));
        CU(cudaMemsetAsync(w->tstarts[i], 0, (size_t)w->ntiles * w->tss * sizeof(uint16_t), w->stream));  // (ghost columns start out empty)
    }
    if (w->strip && !w->comm_stream) {
        // Highest priority: the interior launch has thousands of blocks queued when the exchange is
        // enqueued, and the block scheduler hands freed SM slots to the older grid first -- at equal
        // priority the exchange kernel would only start when the interior has nothing left to dispatch,
        // i.e. not overlap at all.
        int prio_least = 0, prio_greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        CU(cudaStreamCreateWithPriority(&w->comm_stream, cudaStreamNonBlocking, prio_greatest));
        CU(cudaEventCreateWithFlags(&w->ev_edge, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w->ev_exch, cudaEventDisableTiming));
        if (getenv("WRACH_STRIP_EVENTS") == nullptr) {  // (set: the two-launch, event-driven exchange -- A/B runs)
            CU(cudaMalloc(&w->d_sig, 2 * sizeof(uint32_t)));
            CU(cudaMemsetAsync(w->d_sig, 0, 2 * sizeof(uint32_t), w->stream));
        }
    }
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(tile_kernel<WRACH_ARITH_SPV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileS));
        cudaFuncSetAttribute(tile_kernel<WRACH_ARITH_UNFUSED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileS));
        cudaFuncSetAttribute(tile_kernel<WRACH_ARITH_SPV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileS));
        cudaFuncSetAttribute(tile_kernel<WRACH_ARITH_UNFUSED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileS));
    });
    CU(cudaGetLastError());
    return WRACH_OK;
}

float2 *tile_pos(wrach_cuda_worker *w, int buf) { return reinterpret_cast<float2 *>(w->tdata[buf]); }
float2 *tile_vel(wrach_cuda_worker *w, int buf) { return tile_pos(w, buf) + (size_t)w->ntiles * w->tcap; }

// Conversion kernels run over the OWNED tile columns only.
TileConv make_tile_conv(wrach_cuda_worker *w, int buf, uint32_t ord) {
    TileConv c;
    c.gx = w->s.grid_dimensions[0];
    c.gy = w->s.grid_dimensions[1];
    c.ntx = w->ntx;
    c.nty = w->nty;
    c.tcap = w->tcap;
    c.tss = w->tss;
    c.ord = ord;
    c.cells = w->cells;
    c.col_major = (w->strip || w->tile_col_major) ? 1u : 0u;
    c.tx_first = w->t_ghost_l;
    c.x_off = w->t_ghost_l * TileShape::TW;
    c.capacity = w->capacity;
    c.idx = w->idx[w->cur];
    c.pos = w->pos_in;
    c.vel = w->vel_in;
    c.tpos = tile_pos(w, buf);
    c.tvel = tile_vel(w, buf);
    c.ts = w->tstarts[buf];
    c.ctrl = w->ctrl;
    return c;
}
uint32_t owned_tiles(const wrach_cuda_worker *w) { return w->t_own_tc * w->nty; }

// tiles -> the reference's packed layout (no frame may be pending)
int make_packed(wrach_cuda_worker *w) {
    if (w->packed_valid) return WRACH_OK;
    if (!w->tiled_valid) return fail(w, WRACH_ERR_STATE, "no valid copy of the state");
    if (!w->slow_ticket) CU(cudaMalloc(&w->slow_ticket, sizeof(uint32_t)));
    const TileConv c = make_tile_conv(w, w->tcur, 0);
    CU(cudaMemsetAsync(w->slow_ticket, 0, sizeof(uint32_t), w->stream));
    k_tile_pack_counts<TileShape::TW, TileShape::TH><<<owned_tiles(w), 256, 0, w->stream>>>(c);
    k_slow_scan<<<(w->total_cells + 1023) / 1024, 256, 0, w->stream>>>(c.idx, w->total_cells, w->tile_status, ++w->epoch,
                                                                     w->slow_ticket);
    k_tile_pack_copy<TileShape::TW, TileShape::TH><<<owned_tiles(w), 256, 0, w->stream>>>(c);
    w->stats.kernel_launches += 3;
    w->stats.tile_packs++;
    CU(cudaGetLastError());
    w->packed_valid = true;
    w->ckpt_at = w->stats.steps_completed;
    if (w->strip) {  // the strip's population may have changed: learn it (and whether it still fits) now
        CU(cudaMemcpyAsync(w->h_ctrl, w->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, w->stream));
        CU(cudaMemcpyAsync(&w->h_count, c.idx + w->cells + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        if (w->h_ctrl->strip_error == 2u) {
            fail(w, WRACH_ERR_CAPACITY, "strip %d grew past its %u particle slots (arrivals from the neighbouring strips): "
                 "create strips with head-room", w->rank, w->capacity);
            return die(w, WRACH_ERR_CAPACITY);
        }
        w->s.particles_in_frame_count = w->h_count;
    }
    return WRACH_OK;
}

// Stream memory operations (wait for / write a 32-bit word from a stream), fetched from the driver
// through the runtime: they let the exchange stream follow the kernels' own progress counters.
struct StreamMemOps {
    CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
};
StreamMemOps *stream_memops() {
    static StreamMemOps ops;
    static std::once_flag once;
    std::call_once(once, [] {
        void *a = nullptr, *b = nullptr;
        cudaDriverEntryPointQueryResult qa, qb;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &a, cudaEnableDefault, &qa) == cudaSuccess && qa == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuStreamWriteValue32", &b, cudaEnableDefault, &qb) == cudaSuccess && qb == cudaDriverEntryPointSuccess) {
            ops.wait32 = reinterpret_cast<decltype(ops.wait32)>(a);
            ops.write32 = reinterpret_cast<decltype(ops.write32)>(b);
        }
        cudaGetLastError();
    });
    return ops.wait32 && ops.write32 ? &ops : nullptr;
}

// Virtual-memory-management entry points of the driver (shareable allocations), fetched through the
// runtime like the stream memory operations above: the library does not link libcuda.
struct VmmApi {
    CUresult (*GetGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*AddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*Export)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*Import)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
};
VmmApi *vmm_api() {
    static VmmApi api;
    static bool ok = false;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char *name, void **out) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, out, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *out;
        };
        ok = get("cuMemGetAllocationGranularity", reinterpret_cast<void **>(&api.GetGranularity)) &&
             get("cuMemCreate", reinterpret_cast<void **>(&api.Create)) && get("cuMemRelease", reinterpret_cast<void **>(&api.Release)) &&
             get("cuMemAddressReserve", reinterpret_cast<void **>(&api.AddressReserve)) &&
             get("cuMemAddressFree", reinterpret_cast<void **>(&api.AddressFree)) && get("cuMemMap", reinterpret_cast<void **>(&api.Map)) &&
             get("cuMemUnmap", reinterpret_cast<void **>(&api.Unmap)) && get("cuMemSetAccess", reinterpret_cast<void **>(&api.SetAccess)) &&
             get("cuMemExportToShareableHandle", reinterpret_cast<void **>(&api.Export)) &&
             get("cuMemImportFromShareableHandle", reinterpret_cast<void **>(&api.Import));
        cudaGetLastError();
    });
    return ok ? &api : nullptr;
}
CUmemAllocationProp shareable_prop(int device) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}
// map `handle` (bytes, a multiple of the granularity) read-write for `device`; returns 0 on failure
CUdeviceptr vmm_map(VmmApi *vm, CUmemGenericAllocationHandle handle, size_t bytes, int device) {
    CUdeviceptr va = 0;
    if (vm->AddressReserve(&va, bytes, 0, 0, 0) != CUDA_SUCCESS) return 0;
    if (vm->Map(va, bytes, 0, handle, 0) != CUDA_SUCCESS) {
        vm->AddressFree(va, bytes);
        return 0;
    }
    CUmemAccessDesc access = {};
    access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    access.location.id = device;
    access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (vm->SetAccess(va, bytes, &access, 1) != CUDA_SUCCESS) {
        vm->Unmap(va, bytes);
        vm->AddressFree(va, bytes);
        return 0;
    }
    return va;
}
void free_particle_buffer(wrach_cuda_worker *w, int which, float2 *p) {  // cudaMalloc'ed, or a shareable mapping
    wrach_cuda_worker::Shared &sh = w->shared[which];
    if (sh.va && reinterpret_cast<float2 *>(sh.va) == p) {
        if (VmmApi *vm = vmm_api()) {
            vm->Unmap(sh.va, sh.bytes);
            vm->AddressFree(sh.va, sh.bytes);
            vm->Release(sh.handle);
        }
        sh = wrach_cuda_worker::Shared{};
    } else {
        cudaFree(p);
    }
}

struct ColRange {
    uint32_t first, count;
};
// One launch over up to three ranges of tile columns, in that order.
void launch_tile_frame(wrach_cuda_worker *w, const TileFrame &tf0, ColRange a, ColRange b = {0, 0}, ColRange c = {0, 0}) {
    if (a.count == 0) {  // keep the non-empty ranges in front
        a = b;
        b = c;
        c = {0, 0};
    }
    if (a.count == 0) {
        a = b;
        b = {0, 0};
    }
    if (b.count == 0) {
        b = c;
        c = {0, 0};
    }
    if (a.count + b.count + c.count == 0) return;
    TileFrame tf = tf0;
    tf.tx_first = a.first;
    tf.n_first = a.count;
    tf.tx_second = b.first;
    tf.n_second = b.count;
    tf.tx_third = c.first;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.count + b.count + c.count) * w->nty);
    cfg.blockDim = dim3(TileShape::NT);
    cfg.dynamicSmemBytes = sizeof(TileS);
    cfg.stream = w->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tf.pdl ? 1 : 0;
    const bool strip = tf.col_major != 0u;
    if (w->arith == WRACH_ARITH_SPV) {
        if (strip) cudaLaunchKernelEx(&cfg, tile_kernel<WRACH_ARITH_SPV, true>, tf);
        else cudaLaunchKernelEx(&cfg, tile_kernel<WRACH_ARITH_SPV, false>, tf);
    } else {
        if (strip) cudaLaunchKernelEx(&cfg, tile_kernel<WRACH_ARITH_UNFUSED, true>, tf);
        else cudaLaunchKernelEx(&cfg, tile_kernel<WRACH_ARITH_UNFUSED, false>, tf);
    }
    w->stats.kernel_launches++;
}

// Strip workers, NCCL mode: hand the two edge tile columns of buffer `buf` to the neighbouring strips
// and take theirs into the ghost columns (a tile column is one contiguous range of each array).
int tile_exchange_nccl(wrach_cuda_worker *w, int buf, cudaStream_t stream) {
    if (!w->edge_mask || !w->comm) return WRACH_OK;
    NcclApi *nc = nccl_api();
    const size_t col_slots = (size_t)w->nty * w->tcap, col_ts = (size_t)w->nty * w->tss;
    NC(nc->GroupStart());
    for (int side = 0; side < 2; side++) {
        if (!((w->edge_mask >> side) & 1u)) continue;
        const int other = w->rank + (side == 0 ? -1 : 1);
        const size_t edge = side == 0 ? w->t_ghost_l : w->t_ghost_l + w->t_own_tc - 1;  // owned column next to that neighbour
        const size_t ghost = side == 0 ? 0 : w->ntx - 1;
        NC(nc->Send(tile_pos(w, buf) + edge * col_slots, col_slots * sizeof(float2), ncclUint8, other, w->comm, stream));
        NC(nc->Send(tile_vel(w, buf) + edge * col_slots, col_slots * sizeof(float2), ncclUint8, other, w->comm, stream));
        NC(nc->Send(w->tstarts[buf] + edge * col_ts, col_ts * sizeof(uint16_t), ncclUint8, other, w->comm, stream));
        NC(nc->Recv(tile_pos(w, buf) + ghost * col_slots, col_slots * sizeof(float2), ncclUint8, other, w->comm, stream));
        NC(nc->Recv(tile_vel(w, buf) + ghost * col_slots, col_slots * sizeof(float2), ncclUint8, other, w->comm, stream));
        NC(nc->Recv(w->tstarts[buf] + ghost * col_ts, col_ts * sizeof(uint16_t), ncclUint8, other, w->comm, stream));
        w->stats.halo_bytes_sent += 2 * col_slots * sizeof(float2) + col_ts * sizeof(uint16_t);
    }
    NC(nc->GroupEnd());
    return WRACH_OK;
}

// ... in-process mode: copy the neighbours' edge columns into this worker's ghost columns
int tile_exchange_peers(wrach_cuda_worker *w) {
    const size_t col_slots = (size_t)w->nty * w->tcap, col_ts = (size_t)w->nty * w->tss;
    for (int side = 0; side < 2; side++) {
        wrach_cuda_worker *p = w->peer[side];
        if (!p) continue;
        const size_t ghost = side == 0 ? 0 : w->ntx - 1;
        const size_t edge = side == 0 ? p->t_ghost_l + p->t_own_tc - 1 : p->t_ghost_l;  // the neighbour's column facing us
        CU(cudaMemcpyAsync(tile_pos(w, w->tcur) + ghost * col_slots, tile_pos(p, p->tcur) + edge * col_slots, col_slots * sizeof(float2), cudaMemcpyDefault, w->stream));
        CU(cudaMemcpyAsync(tile_vel(w, w->tcur) + ghost * col_slots, tile_vel(p, p->tcur) + edge * col_slots, col_slots * sizeof(float2), cudaMemcpyDefault, w->stream));
        CU(cudaMemcpyAsync(w->tstarts[w->tcur] + ghost * col_ts, p->tstarts[p->tcur] + edge * col_ts, col_ts * sizeof(uint16_t), cudaMemcpyDefault, w->stream));
        w->stats.halo_bytes_sent += 2 * col_slots * sizeof(float2) + col_ts * sizeof(uint16_t);
    }
    return WRACH_OK;
}

TileFrame make_tile_frame(wrach_cuda_worker *w) {
    TileFrame tf;
    tf.lim = make_limits(w->s);
    tf.gx = w->strip ? w->t_gx : w->s.grid_dimensions[0];
    tf.gy = w->s.grid_dimensions[1];
    tf.ntx = w->ntx;
    tf.nty = w->nty;
    tf.tcap = w->tcap;
    tf.tss = w->tss;
    tf.ord = w->tile_ord++;
    tf.pdl = w->pdl_active ? 1u : 0u;
    tf.col_major = (w->strip || w->tile_col_major) ? 1u : 0u;
    tf.tx_first = 0;
    tf.n_first = 0xFFFFFFFFu;
    tf.tx_second = 0;
    tf.n_second = 0;
    tf.tx_third = 0;
    tf.n_edge_blocks = 0;
    tf.ghost_target = 0;
    tf.ghost_ready = nullptr;
    tf.edge_done = nullptr;
    tf.col0 = w->strip ? (int32_t)w->col0 - (int32_t)(w->t_ghost_l * TileShape::TW) : 0;
    tf.in_pos = tile_pos(w, w->tcur);
    tf.in_vel = tile_vel(w, w->tcur);
    tf.out_pos = tile_pos(w, w->tcur ^ 1);
    tf.out_vel = tile_vel(w, w->tcur ^ 1);
    tf.ts_in = w->tstarts[w->tcur];
    tf.ts_out = w->tstarts[w->tcur ^ 1];
    tf.ctrl = w->ctrl;
    return tf;
}

// packed -> tiles, after an upload (or after frames on the other path)
int tile_unpack(wrach_cuda_worker *w) {
    const TileConv c = make_tile_conv(w, w->tcur, w->tile_ord);
    k_tile_unpack<TileShape::TW, TileShape::TH><<<owned_tiles(w), 256, 0, w->stream>>>(c);
    w->stats.kernel_launches++;
    w->stats.tile_unpacks++;
    w->tiled_valid = true;
    CU(cudaGetLastError());
    return WRACH_OK;
}

int enqueue_tile_frames(wrach_cuda_worker *w, uint64_t n, bool profile, float *phys_ms) {
    if (w->tile_pending == 0) {
        w->tile_first_ord = w->tile_ord;
        w->tile_first_from_packed = !w->tiled_valid;
    }
    const bool nccl_strip = w->strip && w->comm && w->edge_mask;
    if (!w->tiled_valid) {
        w->ckpt_at = w->stats.steps_completed;  // nothing is pending and the packed copy is current: the checkpoint
        int rc = tile_unpack(w);
        if (rc) return rc;
        if (nccl_strip) {
            // every strip takes the same path: a scene one of them cannot hold in tiles sends all of them
            // to k_phys / k_rebin (the flag makes the frames below no-ops everywhere), then the first ghosts
            NcclApi *nc = nccl_api();
            NC(nc->AllReduce(&w->ctrl->tile_fail, &w->ctrl->tile_fail, 1, ncclUint32, ncclMax, w->comm, w->stream));
            rc = tile_exchange_nccl(w, w->tcur, w->stream);
            if (rc) return die(w, rc);
            if (StreamMemOps *mo = w->d_sig ? stream_memops() : nullptr) {
                if (mo->write32((CUstream)w->stream, (CUdeviceptr)(w->d_sig + 1), ++w->ghost_seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
                    return die(w, fail(w, WRACH_ERR_CUDA, "cuStreamWriteValue32 failed"));
            }
            CU(cudaEventRecord(w->ev_exch, w->stream));
        }
    }
    if (w->tile_pending == 0) w->tile_first_buf = w->tcur;
    // (programmatic dependent launch between consecutive tile frames pays at every size: one launch per
    // frame, so even a world of a single wave of blocks hides its launch gap -- 1 M: 28.9 -> 27.8 us/frame;
    // the three-launch packed path keeps its threshold, make_frame)
    w->pdl_active = w->pdl;
    for (uint64_t i = 0; i < n; i++) {
        const TileFrame tf = make_tile_frame(w);
        if (profile) CU(cudaEventRecord(w->ev[1], w->stream));
        if (nccl_strip) {
            // The tile columns next to a neighbouring strip come first in the launch (they read the ghosts
            // of the input buffer and produce what the neighbours need next); the ghost exchange of the
            // OUTPUT buffer runs on a second stream as soon as those blocks are done, beside the interior.
            const uint32_t first = w->t_ghost_l, last = w->t_ghost_l + w->t_own_tc - 1;
            const bool el = (w->edge_mask & 1u) != 0, er = (w->edge_mask & 2u) != 0 && (last != first || !el);
            const uint32_t lo = first + (el ? 1u : 0u), hi = last + 1u - (er ? 1u : 0u);
            const ColRange r_el = {first, el ? 1u : 0u}, r_er = {last, er ? 1u : 0u}, r_in = {lo, hi > lo ? hi - lo : 0u};
            StreamMemOps *mo = w->d_sig ? stream_memops() : nullptr;
            if (mo) {
                // one launch: the edge blocks spin until the previous exchange has been delivered and count
                // themselves done; the exchange stream waits for that count (stream memory operations)
                TileFrame t2 = tf;
                t2.n_edge_blocks = (r_el.count + r_er.count) * w->nty;
                t2.ghost_target = w->ghost_seq;
                t2.ghost_ready = w->d_sig + 1;
                t2.edge_done = w->d_sig;
                launch_tile_frame(w, t2, r_el, r_er, r_in);
                w->edge_cum += t2.n_edge_blocks;
                if (mo->wait32((CUstream)w->comm_stream, (CUdeviceptr)w->d_sig, w->edge_cum, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                    return die(w, fail(w, WRACH_ERR_CUDA, "cuStreamWaitValue32 failed"));
                int rc = tile_exchange_nccl(w, w->tcur ^ 1, w->comm_stream);
                if (rc) return die(w, rc);
                if (mo->write32((CUstream)w->comm_stream, (CUdeviceptr)(w->d_sig + 1), ++w->ghost_seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
                    return die(w, fail(w, WRACH_ERR_CUDA, "cuStreamWriteValue32 failed"));
                CU(cudaEventRecord(w->ev_exch, w->comm_stream));
            } else {
                CU(cudaStreamWaitEvent(w->stream, w->ev_exch, 0));  // the ghosts of the input buffer have arrived
                launch_tile_frame(w, tf, r_el, r_er);
                CU(cudaEventRecord(w->ev_edge, w->stream));
                launch_tile_frame(w, tf, r_in);
                CU(cudaStreamWaitEvent(w->comm_stream, w->ev_edge, 0));
                int rc = tile_exchange_nccl(w, w->tcur ^ 1, w->comm_stream);
                if (rc) return die(w, rc);
                CU(cudaEventRecord(w->ev_exch, w->comm_stream));
            }
        } else {
            launch_tile_frame(w, tf, {w->t_ghost_l, w->t_own_tc});
        }
        w->tcur ^= 1;
        w->tile_pending += 1;
        w->packed_valid = false;
        if (profile) {
            CU(cudaEventRecord(w->ev[2], w->stream));
            CU(cudaEventSynchronize(w->ev[2]));
            float a = 0;
            CU(cudaEventElapsedTime(&a, w->ev[1], w->ev[2]));
            *phys_ms += a;
        }
    }
    if (nccl_strip) {
        CU(cudaStreamWaitEvent(w->stream, w->ev_exch, 0));  // a drained main stream means the ghosts are home too
        // The batch ends with a vote: the earliest frame ANY strip's tiles could not hold.  A strip that
        // failed has skipped its later frames and the others have computed on from stale ghosts, so all
        // of them go back together (resolve_tiles).
        if (!w->tile_vote) {
            CU(cudaMalloc(&w->tile_vote, 2 * sizeof(uint32_t)));
            CU(cudaMallocHost(&w->h_tile_vote, 2 * sizeof(uint32_t)));
        }
        NcclApi *nc = nccl_api();
        k_tile_vote<<<1, 1, 0, w->stream>>>(w->ctrl, w->tile_vote);
        w->stats.kernel_launches++;
        NC(nc->AllReduce(w->tile_vote, w->tile_vote, 2, ncclUint32, ncclMax, w->comm, w->stream));
        CU(cudaMemcpyAsync(w->h_tile_vote, w->tile_vote, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    }
    CU(cudaMemcpyAsync(w->h_ctrl, w->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaGetLastError());
    return WRACH_OK;
}

// After a stream synchronisation with tile frames pending: account for them; if the tiles could not
// hold one of them, go back to its input, pack it, and replay from there on k_phys / k_rebin.
int resolve_tiles(wrach_cuda_worker *w) {
    const uint64_t n = w->tile_pending;
    w->tile_pending = 0;
    uint32_t failed = w->h_ctrl->tile_fail, why = w->h_ctrl->tile_why;
    const bool voted = w->strip && w->comm && w->edge_mask;
    if (voted) {  // what all the strips agreed on (enqueue_tile_frames)
        failed = w->h_tile_vote[0] ? ~w->h_tile_vote[0] : 0u;
        why = w->h_tile_vote[0] ? 3u - w->h_tile_vote[1] : 0u;
    }
    if (!failed) {
        w->stats.steps_completed += n;
        w->stats.tile_frames += n;
        return WRACH_OK;
    }
    const uint64_t k = (uint32_t)(failed - 1u - w->tile_first_ord);  // frames that completed before it
    if (k > n) return fail(w, WRACH_ERR_STATE, "tile frame %u failed outside the pending batch", failed - 1u);
    if (w->strip && !voted) return fail(w, WRACH_ERR_STATE, "in-process strips are stepped with wrach_cuda_strip_group_step");
    w->stats.steps_completed += k;
    w->stats.tile_frames += k;
    w->stats.tile_fallbacks++;
    if (why == kTileWhyFar) w->tiles_retry_at = w->stats.steps_completed + 8;
    else w->tiles_off_until_upload = true;
    w->h_ctrl->tile_fail = 0;
    CU(cudaMemsetAsync(&w->ctrl->tile_fail, 0, 2 * sizeof(uint32_t), w->stream));  // tile_fail + tile_why
    if (k == 0 && w->tile_first_from_packed) {  // the unpack (or the very first frame) failed: the packed state is still current
        w->packed_valid = true;
        w->tiled_valid = false;
    } else if (voted) {
        // Strips: the strip that failed skipped its later frames, the others went on with stale ghosts --
        // nobody but the failed strip still holds the input of frame k.  The packed buffers do hold the
        // state of the last upload / read-back (tile frames never touch them): every strip unpacks that
        // checkpoint and replays the frames up to k, which succeeded the first time and are deterministic.
        if (w->stats.steps_completed < w->ckpt_at) return die(w, fail(w, WRACH_ERR_STATE, "strip %d: no checkpoint before the failed frame", w->rank));
        const uint64_t replay = w->stats.steps_completed - w->ckpt_at;
        w->packed_valid = true;
        w->tiled_valid = false;
        if (replay) {
            const uint64_t retry_at = w->tiles_retry_at;
            const bool off = w->tiles_off_until_upload;
            float ms = 0;
            w->stats.steps_completed = w->ckpt_at;  // (enqueue_tile_frames notes the checkpoint it unpacks)
            int rc = enqueue_tile_frames(w, replay, false, &ms);
            if (rc) return die(w, rc);
            CU(cudaStreamSynchronize(w->stream));
            w->tile_pending = 0;
            w->stats.steps_completed += replay;
            w->tiles_retry_at = retry_at;
            w->tiles_off_until_upload = off;
            if (w->h_tile_vote[0]) return die(w, fail(w, WRACH_ERR_STATE, "strip %d: a replayed tile frame failed", w->rank));
            int rc2 = make_packed(w);
            if (rc2) return rc2;
        }
    } else {
        w->tcur = w->tile_first_buf ^ (int)(k & 1u);
        w->tiled_valid = true;
        w->packed_valid = false;
        int rc = make_packed(w);
        if (rc) return rc;
    }
    return n > k ? enqueue_frames(w, n - k, false, nullptr, nullptr) : WRACH_OK;
}

// builder.rs:86-89, once per frame; no host synchronisation.  Every frame is accounted for (pending,
// idx role) as soon as it has been launched, so an error return mid-batch leaves the books right.
// The batch ends with an asynchronous copy of the control block to its pinned mirror: resolve() then
// needs ONE stream synchronisation and no further round trip to learn how the frames went.
int enqueue_frames(wrach_cuda_worker *w, uint64_t n, bool profile, float *phys_ms, float *rebin_ms) {
    if (w->dead) return fail(w, WRACH_ERR_STATE, "worker unusable after an earlier fatal error: %s", w->dead_why.c_str());
    if (n && w->strip && w->edge_mask && !w->comm)
        return fail(w, WRACH_ERR_STATE, "in-process strips are stepped with wrach_cuda_strip_group_step");
    if (n && tiles_usable(w)) {
        if (w->pending) {  // frames of the other kind in flight: settle them first (rare: only right after a fallback)
            int rc = resolve(w);
            if (rc) return rc;
        }
        int rc = tiles_allocate(w);
        if (rc) return rc;
        if (tiles_usable(w)) {
            float ms = 0;
            rc = enqueue_tile_frames(w, n, profile, &ms);
            if (profile) *phys_ms += ms;
            return rc;
        }
    }
    if (n) {
        if (w->tile_pending) {
            int rc = resolve(w);
            if (rc) return rc;
        }
        int rc = make_packed(w);
        if (rc) return rc;
        w->tiled_valid = false;
    }
    for (uint64_t i = 0; i < n; i++) {
        if (w->strip && w->edge_mask && !w->comm)
            return fail(w, WRACH_ERR_STATE, "in-process strips are stepped with wrach_cuda_strip_group_step");
        const Frame f = make_frame(w, w->cur_enqueue);
        if (profile) CU(cudaEventRecord(w->ev[1], w->stream));
        if (w->neighbour_mode) {
            if (w->strip && w->edge_mask) {
                launch_nb_halo_pack(w, f);
                int rc = nb_halo_exchange_nccl(w);
                if (rc) return die(w, rc);
            }
            launch_neighbours(w, f);
        }
        launch_phys(w, f);
        w->cur_enqueue ^= 1;  // from here on the frame exists: account for it whatever happens next
        w->pending += 1;
        if (profile) CU(cudaEventRecord(w->ev[2], w->stream));
        if (w->strip) {
            int rc = strip_exchange_nccl(w);
            if (rc) return die(w, rc);
        }
        launch_rebin(w, f);
        if (profile) {
            CU(cudaEventRecord(w->ev[3], w->stream));
            CU(cudaEventSynchronize(w->ev[3]));
            float a = 0, b = 0;
            CU(cudaEventElapsedTime(&a, w->ev[1], w->ev[2]));
            CU(cudaEventElapsedTime(&b, w->ev[2], w->ev[3]));
            *phys_ms += a;
            *rebin_ms += b;
        }
    }
    if (n) CU(cudaMemcpyAsync(w->h_ctrl, w->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaGetLastError());
    return WRACH_OK;
}

// Re-bin ONE frame whose physics already ran (pos_out / vel_out valid) with the generic kernels.
int slow_rebin(wrach_cuda_worker *w, int read_role) {
    if (!w->slow_src) {
        CU(cudaMalloc(&w->slow_src, ((size_t)w->capacity + 4) * sizeof(uint32_t)));
        CU(cudaMalloc(&w->slow_cursor, (size_t)w->total_cells * sizeof(uint32_t)));
    }
    if (!w->slow_ticket) CU(cudaMalloc(&w->slow_ticket, sizeof(uint32_t)));  // (make_packed may have made it already)
    Frame f = make_frame(w, read_role);
    const int threads = 256;  // (grid-stride loops over the N the indices hold; the grid is sized for the buffers)
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(((uint64_t)w->capacity + threads) / threads, 148u * 16u);
    CU(cudaMemsetAsync(f.starts_next, 0, (size_t)w->total_cells * sizeof(uint32_t), w->stream));
    CU(cudaMemsetAsync(w->run_total, 0, ((size_t)(w->cells + kRun - 1) / kRun + 1) * sizeof(uint32_t), w->stream));
    CU(cudaMemsetAsync(w->slow_cursor, 0, (size_t)w->total_cells * sizeof(uint32_t), w->stream));
    CU(cudaMemsetAsync(w->slow_ticket, 0, sizeof(uint32_t), w->stream));
    k_slow_count<<<blocks, threads, 0, w->stream>>>(f);
    k_slow_scan<<<(w->total_cells + 1023) / 1024, 256, 0, w->stream>>>(f.starts_next, w->total_cells, w->tile_status,
                                                                     f.epoch, w->slow_ticket);
    k_slow_scatter<<<blocks, threads, 0, w->stream>>>(f, w->slow_cursor, w->slow_src);
    k_slow_rank_move<<<blocks, threads, 0, w->stream>>>(f, w->slow_src);
    w->stats.kernel_launches += 4;
    w->stats.slow_path_steps++;
    CU(cudaGetLastError());
    return WRACH_OK;
}

// ---------------------------------------------------------------------------------------------
// Strips: the collective re-bin of one frame whose physics already ran on every strip (pos_out /
// vel_out valid, the fast re-bin skipped everywhere).  Three phases with the count matrix of all
// strips in between; the NCCL driver and the in-process driver share them.
constexpr uint32_t kXrRow = kMaxStrips + 1;  // a strip's gathered row: counts per destination strip, then its capacity

void xr_free(wrach_cuda_worker *w) {
    cudaFree(w->xr_send);
    cudaFree(w->xr_recv);
    cudaFree(w->xr_key);
    w->xr_send = w->xr_recv = nullptr;
    w->xr_key = nullptr;
}

XRebin make_xrebin(wrach_cuda_worker *w, int read_role) {
    XRebin x{};
    x.gs = w->gs;
    x.lgx = w->col1 - w->col0;
    x.col0 = w->col0;
    x.cells = w->cells;
    x.n_ranks = (uint32_t)w->n_ranks;
    x.rank = (uint32_t)w->rank;
    for (int r = 0; r < w->n_ranks; r++) x.col_end[r] = w->col_end[r];
    x.starts = w->idx[read_role];
    x.pos_out = w->pos_out;
    x.vel_out = w->vel_out;
    x.send_cnt = w->xr_cnt;
    x.send_cursor = w->xr_cnt + kMaxStrips;
    x.send = w->xr_send;
    x.recv = w->xr_recv;
    x.starts_next = w->idx[read_role ^ 1];
    x.cursor = w->slow_cursor;
    x.src = w->slow_src;
    x.key_at = w->xr_key;
    x.pos_in = w->pos_in;
    x.vel_in = w->vel_in;
    return x;
}

constexpr uint32_t kXrGrid = 148 * 8;

// phase 1: particles per destination strip -> xr_cnt[0 .. n_ranks), capacity behind them
int xr_phase_count(wrach_cuda_worker *w, int read_role) {
    if (w->n_ranks > kMaxStrips) return fail(w, WRACH_ERR_STATE, "the collective re-bin supports at most %d strips", kMaxStrips);
    if (!w->xr_cnt) {
        CU(cudaMalloc(&w->xr_cnt, (2 * kMaxStrips + 2 + (size_t)kMaxStrips * kXrRow) * sizeof(uint32_t)));
        CU(cudaMallocHost(&w->xr_host, (size_t)kMaxStrips * kXrRow * sizeof(uint32_t)));
    }
    if (!w->slow_src) {
        CU(cudaMalloc(&w->slow_src, ((size_t)w->capacity + 4) * sizeof(uint32_t)));
        CU(cudaMalloc(&w->slow_cursor, (size_t)w->total_cells * sizeof(uint32_t)));
    }
    if (!w->slow_ticket) CU(cudaMalloc(&w->slow_ticket, sizeof(uint32_t)));
    CU(cudaMalloc(&w->xr_send, ((size_t)w->capacity + 1) * sizeof(XRec)));
    CU(cudaMalloc(&w->xr_recv, ((size_t)w->capacity + 1) * sizeof(XRec)));
    CU(cudaMalloc(&w->xr_key, ((size_t)w->capacity + 1) * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(w->xr_cnt, 0, 2 * kMaxStrips * sizeof(uint32_t), w->stream));
    CU(cudaMemcpyAsync(w->xr_cnt + w->n_ranks, &w->capacity, sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));  // row = counts, capacity
    const XRebin x = make_xrebin(w, read_role);
    k_xr_count<<<kXrGrid, 256, 0, w->stream>>>(x);
    // (the capacity word sits in the cursor area: cleared again before k_xr_pack uses it)
    w->stats.kernel_launches++;
    CU(cudaGetLastError());
    return WRACH_OK;
}

// phase 2: with every strip's row (rows[s * stride + d] = particles strip s sends to strip d,
// rows[s * stride + n_ranks] = capacity of strip s): segment offsets, capacity check, the records
int xr_phase_pack(wrach_cuda_worker *w, int read_role, const uint32_t *rows, uint32_t stride, XRebin *out) {
    const uint32_t n = (uint32_t)w->n_ranks;
    XRebin x = make_xrebin(w, read_role);
    // (host logic shared with the CPU tests: every strip checks every strip, so that all of them stop together)
    const int over = wrach_host_strip_exchange_plan(n, (uint32_t)w->rank, rows, stride, x.send_off, nullptr, &x.n_recv);
    if (over >= 0) {
        fail(w, WRACH_ERR_CAPACITY, "strip %d would outgrow the %u particle slots it was created with in this frame: "
             "create strips with head-room", over, rows[(size_t)over * stride + n]);
        return die(w, WRACH_ERR_CAPACITY);
    }
    CU(cudaMemsetAsync(w->xr_cnt + kMaxStrips, 0, kMaxStrips * sizeof(uint32_t), w->stream));
    k_xr_pack<<<kXrGrid, 256, 0, w->stream>>>(x);
    w->stats.kernel_launches++;
    CU(cudaGetLastError());
    *out = x;
    return WRACH_OK;
}

// phase 3: the records have arrived in xr_recv: new indices, canonical order, the frame is done
int xr_phase_place(wrach_cuda_worker *w, const XRebin &x) {
    CU(cudaMemsetAsync(x.starts_next, 0, (size_t)w->total_cells * sizeof(uint32_t), w->stream));
    CU(cudaMemsetAsync(w->run_total, 0, ((size_t)(w->cells + kRun - 1) / kRun + 1) * sizeof(uint32_t), w->stream));
    CU(cudaMemsetAsync(w->slow_cursor, 0, (size_t)w->total_cells * sizeof(uint32_t), w->stream));
    CU(cudaMemsetAsync(w->slow_ticket, 0, sizeof(uint32_t), w->stream));
    for (int i = 0; i < 2; i++)  // the frame's export messages were never consumed (k_import_place clears them on the fast path)
        if (w->exp_buf[i]) CU(cudaMemsetAsync(w->exp_buf[i], 0, 16, w->stream));
    k_xr_cells<<<kXrGrid, 256, 0, w->stream>>>(x);
    k_slow_scan<<<(w->total_cells + 1023) / 1024, 256, 0, w->stream>>>(x.starts_next, w->total_cells, w->tile_status, ++w->epoch,
                                                                     w->slow_ticket);
    k_xr_scatter<<<kXrGrid, 256, 0, w->stream>>>(x);
    k_xr_place<<<kXrGrid, 256, 0, w->stream>>>(x);
    w->stats.kernel_launches += 4;
    w->stats.slow_path_steps++;
    CU(cudaMemsetAsync(&w->ctrl->abort, 0, 2 * sizeof(uint32_t), w->stream));  // abort + far_seen
    CU(cudaGetLastError());
    w->s.particles_in_frame_count = x.n_recv;
    // the frame exists now: same bookkeeping as the single-device generic re-bin
    w->cur ^= 1;
    w->pending -= 1;
    w->stats.steps_completed += 1;
    w->cur_enqueue = w->cur;
    return WRACH_OK;
}

// NCCL mode: every strip calls this from the same resolve()
int strip_collective_rebin_nccl(wrach_cuda_worker *w, int read_role) {
    NcclApi *nc = nccl_api();
    const uint32_t n = (uint32_t)w->n_ranks;
    int rc = xr_phase_count(w, read_role);
    if (rc) return die(w, rc);
    uint32_t *gathered = w->xr_cnt + 2 * kMaxStrips + 2;
    NC(nc->AllGather(w->xr_cnt, gathered, n + 1, ncclUint32, w->comm, w->stream));
    CU(cudaMemcpyAsync(w->xr_host, gathered, (size_t)n * (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    XRebin x;
    rc = xr_phase_pack(w, read_role, w->xr_host, n + 1, &x);
    if (rc) return rc;
    const uint32_t *rows = w->xr_host;
    uint32_t recv_off = 0;
    NC(nc->GroupStart());
    for (uint32_t r = 0; r < n; r++) {
        const uint32_t out = rows[(uint32_t)w->rank * (n + 1) + r], in = rows[r * (n + 1) + (uint32_t)w->rank];
        if (r == (uint32_t)w->rank) {
            if (in) CU(cudaMemcpyAsync(x.recv + recv_off, x.send + x.send_off[r], (size_t)in * sizeof(XRec), cudaMemcpyDeviceToDevice, w->stream));
        } else {
            if (out) NC(nc->Send(x.send + x.send_off[r], (size_t)out * sizeof(XRec), ncclUint8, (int)r, w->comm, w->stream));
            if (in) NC(nc->Recv(x.recv + recv_off, (size_t)in * sizeof(XRec), ncclUint8, (int)r, w->comm, w->stream));
            w->stats.halo_bytes_sent += (size_t)out * sizeof(XRec);
        }
        recv_off += in;
    }
    NC(nc->GroupEnd());
    rc = xr_phase_place(w, x);
    if (rc) return die(w, rc);
    CU(cudaStreamSynchronize(w->stream));
    xr_free(w);
    return WRACH_OK;
}

// In-process mode: the same three phases over all the workers of the group.  Every worker has run the
// frame's physics (and nothing after it) and carries the far flag.
int strip_collective_rebin_group(wrach_cuda_worker **workers, int n) {
    auto on = [&](int i) { cudaSetDevice(workers[i]->device); return workers[i]; };
    std::vector<uint32_t> rows((size_t)n * (n + 1));
    for (int i = 0; i < n; i++) {
        wrach_cuda_worker *w = on(i);
        int rc = xr_phase_count(w, w->cur);
        if (rc) return die(w, rc);
        CU(cudaMemcpyAsync(w->xr_host, w->xr_cnt, (size_t)(n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    }
    for (int i = 0; i < n; i++) {
        wrach_cuda_worker *w = on(i);
        CU(cudaStreamSynchronize(w->stream));
        for (int d = 0; d <= n; d++) rows[(size_t)i * (n + 1) + d] = w->xr_host[d];
    }
    std::vector<XRebin> xs(n);
    for (int i = 0; i < n; i++) {
        wrach_cuda_worker *w = on(i);
        int rc = xr_phase_pack(w, w->cur, rows.data(), (uint32_t)n + 1, &xs[i]);
        if (rc) {
            if (w != workers[0]) workers[0]->err = w->err;
            for (int j = 0; j < n; j++) die(workers[j], rc);
            return rc;
        }
    }
    for (int i = 0; i < n; i++) {
        wrach_cuda_worker *w = on(i);
        CU(cudaStreamSynchronize(w->stream));
    }
    for (int i = 0; i < n; i++) {  // worker i fetches its segment from every worker's send buffer
        wrach_cuda_worker *w = on(i);
        uint32_t recv_off = 0;
        for (int s_ = 0; s_ < n; s_++) {
            const uint32_t cnt = rows[(size_t)s_ * (n + 1) + i];
            if (cnt) CU(cudaMemcpyAsync(xs[i].recv + recv_off, xs[s_].send + xs[s_].send_off[i], (size_t)cnt * sizeof(XRec), cudaMemcpyDefault, w->stream));
            if (s_ != i) w->stats.halo_bytes_sent += (size_t)cnt * sizeof(XRec);
            recv_off += cnt;
        }
        int rc = xr_phase_place(w, xs[i]);
        if (rc) return die(w, rc);
    }
    for (int i = 0; i < n; i++) {
        wrach_cuda_worker *w = on(i);
        CU(cudaStreamSynchronize(w->stream));
    }
    for (int i = 0; i < n; i++) xr_free(on(i));
    return WRACH_OK;
}

// Wait for the enqueued frames; if a frame hit the far-mover flag, finish it on the generic path
// and re-enqueue what was skipped behind it.
int resolve(wrach_cuda_worker *w) {
    while (true) {
        CU(cudaStreamSynchronize(w->stream));  // also completes the control-block mirror enqueue_frames queued
        if (w->tile_pending) {
            int rc = resolve_tiles(w);  // (a fallback re-enqueues on the other path: go round again)
            if (rc) return rc;
            continue;
        }
        if (w->pending == 0) return WRACH_OK;
        const uint32_t completed = w->h_ctrl->steps_done - w->steps_done_seen;
        w->steps_done_seen = w->h_ctrl->steps_done;
        w->cur ^= (int)(completed & 1u);
        w->pending -= completed;
        w->stats.steps_completed += completed;
        if (w->h_ctrl->dense_seen) w->dense_enabled = true;
        if (w->h_ctrl->strip_error == 2u) {
            fail(w, WRACH_ERR_CAPACITY, "strip %d grew past its %u particle slots (arrivals from the neighbouring strips): "
                 "create strips with head-room", w->rank, w->capacity);
            return die(w, WRACH_ERR_CAPACITY);
        }
        if (w->h_ctrl->strip_error) {
            fail(w, WRACH_ERR_NCCL, "strip %d received a malformed exchange message", w->rank);
            return die(w, WRACH_ERR_NCCL);
        }
        if (w->strip && w->edge_mask && (w->h_ctrl->abort || w->h_ctrl->far_seen)) {
            // a far mover (or more leavers than a message holds) on SOME strip: the flag was reduced over
            // all of them before any re-bin kernel ran, so every strip is here with the same frame pending
            if (w->pending == 0) return fail(w, WRACH_ERR_STATE, "abort flag set with no frame pending");
            if (!w->comm) return fail(w, WRACH_ERR_STATE, "in-process strips are stepped with wrach_cuda_strip_group_step");
            int rc = strip_collective_rebin_nccl(w, w->cur);
            if (rc) return die(w, rc);  // (half a collective step: the strips no longer agree on anything)
        } else if (w->h_ctrl->abort || w->h_ctrl->far_seen) {
            if (w->pending == 0) return fail(w, WRACH_ERR_STATE, "abort flag set with no frame pending");
            int rc = slow_rebin(w, w->cur);
            if (rc) return rc;
            CU(cudaMemsetAsync(&w->ctrl->abort, 0, 2 * sizeof(uint32_t), w->stream));  // abort + far_seen
            w->cur ^= 1;
            w->pending -= 1;
            w->stats.steps_completed += 1;
        } else if (w->pending != 0) {
            return fail(w, WRACH_ERR_STATE, "%llu frames unaccounted for", (unsigned long long)w->pending);
        }
        w->cur_enqueue = w->cur;
        if (w->pending) {
            const uint64_t again = w->pending;
            w->pending = 0;
            int rc = enqueue_frames(w, again, false, nullptr, nullptr);
            if (rc) return rc;
        }
    }
}

// Wait for whatever is in flight and make the reference's packed layout current (what every
// host-visible buffer access sees).
int settle_packed(wrach_cuda_worker *w) {
    if (w->pending || w->tile_pending) {
        int rc = resolve(w);
        if (rc) return rc;
    }
    return make_packed(w);
}

void *buffer_ptr(wrach_cuda_worker *w, wrach_buffer b, size_t *bytes) {
    const size_t pb = (size_t)w->capacity * sizeof(float2), ib = (size_t)w->total_cells * sizeof(uint32_t);
    switch (b) {
        case WRACH_INDICES_MAIN: *bytes = ib; return w->idx[w->cur];
        case WRACH_INDICES_BLOCK_SUMS: *bytes = ib; return w->idx[w->cur ^ 1];
        case WRACH_POSITIONS_IN: *bytes = pb; return w->pos_in;
        case WRACH_POSITIONS_OUT: *bytes = pb; return w->pos_out;
        case WRACH_VELOCITIES_IN: *bytes = pb; return w->vel_in;
        case WRACH_VELOCITIES_OUT: *bytes = pb; return w->vel_out;
        default: *bytes = 0; return nullptr;
    }
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int create_common(wrach_cuda_worker *w) {
    DeviceGuard guard(w->device);  // the caller's current device is restored on return
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, w->device));
    if (prop.major < 10)
        return fail(w, WRACH_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", w->device,
                    prop.major, prop.minor);
    w->dense_grid = (uint32_t)prop.multiProcessorCount * kDenseBlocksPerSM;
    w->resident_phys_blocks = (uint32_t)prop.multiProcessorCount * (uint32_t)WRACH_PHYS_MINBLOCKS;
    if (const char *e = getenv("WRACH_PDL")) {
        w->pdl = e[0] != '0';
        w->pdl_forced = e[0] == '2';  // also on worlds of a single wave of blocks (A/B runs)
    }
    if (const char *e = getenv("WRACH_TILES")) w->tiles_on = e[0] != '0';
    if (const char *e = getenv("WRACH_TILE_COLMAJOR")) w->tile_col_major = e[0] == '1';  // tuning: the strips' tile order on one device
    CU(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    const size_t pb = ((size_t)w->capacity + 4) * sizeof(float2), ib = ((size_t)w->total_cells + 4) * sizeof(uint32_t);
    for (int i = 0; i < 2; i++) {
        CU(cudaMalloc(&w->idx[i], ib));
        CU(cudaMemsetAsync(w->idx[i], 0, ib, w->stream));
    }
    float2 **bufs[4] = {&w->pos_in, &w->vel_in, &w->pos_out, &w->vel_out};
    for (auto b : bufs) {
        CU(cudaMalloc(b, pb));
        CU(cudaMemsetAsync(*b, 0, pb, w->stream));  // builder.rs:52-55: zero-filled
    }
    CU(cudaMalloc(&w->meta, ((size_t)w->capacity + 16) * sizeof(uint32_t)));
    CU(cudaMemsetAsync(w->meta, 0, ((size_t)w->capacity + 16) * sizeof(uint32_t), w->stream));
    {
        const size_t runs = ((size_t)w->cells + kRun - 1) / kRun, lists = runs * kVListsPerRun;
        CU(cudaMalloc(&w->cls, ((size_t)w->cells + 16) * sizeof(uint32_t)));
        CU(cudaMemsetAsync(w->cls, 0, ((size_t)w->cells + 16) * sizeof(uint32_t), w->stream));
        // written and read in dense mode / on the general path only
        CU(cudaMalloc(&w->cls9, ((size_t)w->cells + 16) * 9 * sizeof(uint32_t)));
        CU(cudaMalloc(&w->goff9, ((size_t)w->cells + 16) * 9 * sizeof(uint32_t)));
        CU(cudaMalloc(&w->dense_list, (runs * 3 + 1) * 2 * sizeof(uint4)));
        CU(cudaMalloc(&w->run_total, (runs + 8) * sizeof(uint32_t)));  // k_run_scan works in 16-byte groups
        CU(cudaMemsetAsync(w->run_total, 0, (runs + 8) * sizeof(uint32_t), w->stream));
        CU(cudaMalloc(&w->run_base, (runs + 8) * sizeof(uint32_t)));
        CU(cudaMemsetAsync(w->run_base, 0, (runs + 8) * sizeof(uint32_t), w->stream));
        CU(cudaMalloc(&w->vl_slot, lists * kVW * sizeof(uint32_t)));
        CU(cudaMalloc(&w->vl_meta, lists * kVW * sizeof(uint16_t)));
        CU(cudaMalloc(&w->vl_cnt, (lists + 64) * sizeof(uint16_t)));
        CU(cudaMemsetAsync(w->vl_cnt, 0, (lists + 64) * sizeof(uint16_t), w->stream));
    }
    CU(cudaMalloc(&w->ctrl, sizeof(Ctrl)));
    CU(cudaMemsetAsync(w->ctrl, 0, sizeof(Ctrl), w->stream));
    CU(cudaMallocHost(&w->h_ctrl, sizeof(Ctrl)));
    w->n_status = (w->total_cells + 1023) / 1024 + 1;
    CU(cudaMalloc(&w->tile_status, (size_t)w->n_status * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(w->tile_status, 0, (size_t)w->n_status * sizeof(unsigned long long), w->stream));
    for (auto &e : w->ev) CU(cudaEventCreate(&e));
    CU(cudaStreamSynchronize(w->stream));
    return WRACH_OK;
}

}  // namespace

extern "C" {

const char *wrach_cuda_version(void) { return "wrach_cuda sm_100a r2"; }

int wrach_cuda_create(const wrach_world_settings *settings, uint32_t total_cells, uint32_t max_particles,
                      int device, int arith, wrach_cuda_worker **out) {
    if (!settings || !out) return fail(nullptr, WRACH_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (arith != WRACH_ARITH_UNFUSED && arith != WRACH_ARITH_SPV)
        return fail(nullptr, WRACH_ERR_BAD_ARG, "unknown arithmetic variant %d", arith);
    int rc = validate_settings(nullptr, *settings, total_cells, max_particles);
    if (rc) return rc;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(nullptr, WRACH_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= count) return fail(nullptr, WRACH_ERR_BAD_ARG, "device %d out of range", device);
    wrach_cuda_worker *w = new (std::nothrow) wrach_cuda_worker();
    if (!w) return fail(nullptr, WRACH_ERR_BAD_ARG, "out of host memory");
    w->device = device;
    w->arith = arith;
    w->s = *settings;
    w->total_cells = total_cells;
    w->cells = total_cells - 2;
    w->capacity = max_particles;
    rc = create_common(w);
    if (rc) {
        g_create_error = w->err;
        wrach_cuda_destroy(w);
        return rc;
    }
    *out = w;
    return WRACH_OK;
}

int wrach_cuda_create_strip(const wrach_world_settings *global_settings, uint32_t max_particles, int device,
                            int arith, int rank, int n_ranks, const void *nccl_unique_id,
                            wrach_cuda_worker **out) {
    if (!global_settings || !out) return fail(nullptr, WRACH_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(nullptr, WRACH_ERR_BAD_ARG, "bad rank %d of %d", rank, n_ranks);
    if (n_ranks > kMaxStrips) return fail(nullptr, WRACH_ERR_BAD_ARG, "at most %d strips are supported", kMaxStrips);
    if (arith != WRACH_ARITH_UNFUSED && arith != WRACH_ARITH_SPV)
        return fail(nullptr, WRACH_ERR_BAD_ARG, "unknown arithmetic variant %d", arith);
    const wrach_world_settings &g = *global_settings;
    const uint64_t gcells = (uint64_t)g.grid_dimensions[0] * g.grid_dimensions[1];
    if (gcells + 2 > 0xFFFFFFFFull) return fail(nullptr, WRACH_ERR_BAD_ARG, "grid too large");
    int rc = validate_settings(nullptr, g, (uint32_t)gcells + 2, 0xFFFFFFFFu);
    if (rc) return rc;
    uint32_t c0, c1;
    wrach_cuda_strip_columns(g.grid_dimensions[0], rank, n_ranks, &c0, &c1);
    if (c1 - c0 < 3) return fail(nullptr, WRACH_ERR_BAD_ARG, "strip %d would own %u columns; at least 3 are needed", rank, c1 - c0);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(nullptr, WRACH_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= count) return fail(nullptr, WRACH_ERR_BAD_ARG, "device %d out of range", device);
    wrach_cuda_worker *w = new (std::nothrow) wrach_cuda_worker();
    if (!w) return fail(nullptr, WRACH_ERR_BAD_ARG, "out of host memory");
    w->device = device;
    w->arith = arith;
    w->strip = true;
    w->rank = rank;
    w->n_ranks = n_ranks;
    w->gs = g;
    w->col0 = c0;
    w->col1 = c1;
    w->edge_mask = (rank > 0 ? 1u : 0u) | (rank + 1 < n_ranks ? 2u : 0u);
    w->s = g;  // global view rectangle, local grid
    w->s.grid_dimensions[0] = c1 - c0;
    w->s.particles_in_frame_count = 0;
    w->cells = (c1 - c0) * g.grid_dimensions[1];
    w->total_cells = w->cells + 2;
    w->capacity = max_particles;
    w->exp_cap = std::max(1024u, 4u * g.grid_dimensions[1]);
    // every strip takes the same path: the tile frames need ALL the cuts on tile boundaries
    w->strip_tiles_ok = true;
    for (int r = 0; r < n_ranks && r < kMaxStrips; r++) {
        uint32_t b = 0, e = 0;
        wrach_cuda_strip_columns(g.grid_dimensions[0], r, n_ranks, &b, &e);
        w->col_end[r] = e;
        if (b % (uint32_t)TileShape::TW != 0) w->strip_tiles_ok = false;
    }
    rc = create_common(w);
    if (!rc && w->edge_mask) rc = [&]() -> int {
        const size_t bytes = msg_bytes(w->exp_cap), rows3 = (size_t)2 * g.grid_dimensions[1] * 3 * sizeof(uint32_t);
        for (int i = 0; i < 2; i++) {
            CU(cudaMalloc(&w->exp_buf[i], bytes));
            CU(cudaMalloc(&w->imp_buf[i], bytes));
            CU(cudaMemset(w->exp_buf[i], 0, bytes));
            CU(cudaMemset(w->imp_buf[i], 0, bytes));
        }
        CU(cudaMalloc(&w->imp_cnt, rows3));
        CU(cudaMalloc(&w->imp_off, rows3));
        CU(cudaMemset(w->imp_cnt, 0, rows3));
        CU(cudaMemset(w->imp_off, 0, rows3));
        if (nccl_unique_id) {
            NcclApi *nc = nccl_api();
            if (!nc) return fail(w, WRACH_ERR_NCCL, "libnccl.so.2 could not be loaded");
            ncclUniqueId id;
            memcpy(&id, nccl_unique_id, sizeof id);
            NC(nc->CommInitRank(&w->comm, n_ranks, id, rank));
        }
        return WRACH_OK;
    }();
    if (rc) {
        g_create_error = w->err;
        wrach_cuda_destroy(w);
        return rc;
    }
    *out = w;
    return WRACH_OK;
}

int wrach_cuda_nccl_unique_id(void *out_128_bytes) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    wrach_cuda_worker *w = nullptr;
    if (!out_128_bytes) return fail(nullptr, WRACH_ERR_BAD_ARG, "null argument");
    NcclApi *nc = nccl_api();
    if (!nc) return fail(nullptr, WRACH_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    NC(nc->GetUniqueId(&id));
    memcpy(out_128_bytes, &id, sizeof id);
    return WRACH_OK;
}

int wrach_cuda_strip_info(const wrach_cuda_worker *w, uint32_t *col_begin, uint32_t *col_end, uint32_t *total_cells) {
    if (!w) return WRACH_ERR_BAD_ARG;
    if (col_begin) *col_begin = w->strip ? w->col0 : 0u;
    if (col_end) *col_end = w->strip ? w->col1 : w->s.grid_dimensions[0];
    if (total_cells) *total_cells = w->total_cells;
    return WRACH_OK;
}

// In-process strips (several workers in one process, on the same or on different devices): every
// frame runs in lockstep -- physics everywhere, then each worker copies its neighbours' exchange
// messages, then the re-bin everywhere.  Same kernels and same results as the NCCL mode; used where
// one process drives all the devices and by the single-GPU tests.
int wrach_cuda_strip_group_step(wrach_cuda_worker **workers, int n, uint32_t n_steps) {
    if (!workers || n < 1 || n > 64) return WRACH_ERR_BAD_ARG;
    for (int i = 0; i < n; i++) {
        wrach_cuda_worker *w = workers[i];
        if (!w || !w->strip || w->rank != i || w->n_ranks != n || w->comm)
            return fail(w, WRACH_ERR_BAD_ARG, "workers must be the in-process strips 0..n-1 of one world, in order");
    }
    // every handle's mutex, in rank order, for the whole call; the caller's device is restored on return
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int i = 0; i < n; i++) locks.emplace_back(workers[i]->mu);
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    struct Restore {
        int dev;
        ~Restore() { if (dev >= 0) cudaSetDevice(dev); }
    } restore{prev_device};
    for (int i = 0; i < n; i++) {
        wrach_cuda_worker *w = workers[i];
        if (w->dead) return fail(w, WRACH_ERR_STATE, "worker unusable after an earlier fatal error: %s", w->dead_why.c_str());
        w->peer[0] = i > 0 ? workers[i - 1] : nullptr;
        w->peer[1] = i + 1 < n ? workers[i + 1] : nullptr;
    }
    auto on = [&](int i) { cudaSetDevice(workers[i]->device); return workers[i]; };
    auto sync_all = [&]() -> int {
        for (int i = 0; i < n; i++) {
            wrach_cuda_worker *w = on(i);
            CU(cudaStreamSynchronize(w->stream));
        }
        return WRACH_OK;
    };
    auto report = [&](wrach_cuda_worker *w, int rc) {  // callers read the message from the first handle
        if (w != workers[0]) workers[0]->err = w->err;
        return rc;
    };
    // Frames run in lockstep, one at a time, on whichever path ALL the strips can take:
    //  * the fused tile frames (columns cut on tile boundaries, a scene that fits): the neighbours' edge
    //    tile columns are copied into the ghost columns before every frame;
    //  * k_phys / k_rebin on the packed layout with the particle exchange otherwise.
    // A frame some strip cannot finish on its path is redone by all of them on the next more general
    // one -- tiles -> packed -> the collective re-bin -- exactly as the NCCL mode does, where the
    // strips learn about each other's trouble through reductions instead of this loop.
    for (uint32_t step = 0; step < n_steps; step++) {
        bool tiles = true;
        for (int i = 0; i < n; i++) tiles = tiles && tiles_usable(workers[i]);
        if (tiles) {
            for (int i = 0; i < n; i++) {
                wrach_cuda_worker *w = on(i);
                if (w->pending || w->tile_pending) {
                    int rc = resolve(w);
                    if (rc) return report(w, rc);
                }
                int rc = tiles_allocate(w);
                if (rc) return report(w, rc);
                if (!tiles_usable(w)) tiles = false;
            }
        }
        if (tiles) {
            bool unpacked = false;
            for (int i = 0; i < n; i++) {
                wrach_cuda_worker *w = on(i);
                if (!w->tiled_valid) {
                    w->ckpt_at = w->stats.steps_completed;
                    int rc = tile_unpack(w);
                    if (rc) return report(w, rc);
                    unpacked = true;
                    CU(cudaMemcpyAsync(w->h_ctrl, w->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, w->stream));
                }
            }
            int rc = sync_all();
            if (rc) return rc;
            if (unpacked) {
                bool failed = false;
                for (int i = 0; i < n; i++) failed = failed || workers[i]->h_ctrl->tile_fail != 0;
                if (failed) {  // the scene does not fit the tiles of some strip: all of them take the other path
                    for (int i = 0; i < n; i++) {
                        wrach_cuda_worker *w = on(i);
                        w->tiles_off_until_upload = true;
                        w->tiled_valid = false;
                        w->stats.tile_fallbacks++;
                        CU(cudaMemsetAsync(&w->ctrl->tile_fail, 0, 2 * sizeof(uint32_t), w->stream));
                    }
                    tiles = false;
                }
            }
        }
        if (tiles) {
            for (int i = 0; i < n; i++) {
                int rc = tile_exchange_peers(on(i));
                if (rc) return report(workers[i], rc);
            }
            int rc = sync_all();
            if (rc) return rc;
            for (int i = 0; i < n; i++) {
                wrach_cuda_worker *w = on(i);
                w->pdl_active = false;
                const TileFrame tf = make_tile_frame(w);
                launch_tile_frame(w, tf, {w->t_ghost_l, w->t_own_tc});
                w->tcur ^= 1;
                w->packed_valid = false;
                CU(cudaMemcpyAsync(w->h_ctrl, w->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, w->stream));
                CU(cudaGetLastError());
            }
            rc = sync_all();
            if (rc) return rc;
            uint32_t why = 0;
            for (int i = 0; i < n; i++)
                if (workers[i]->h_ctrl->tile_fail) why = std::max(why, 3u - workers[i]->h_ctrl->tile_why);  // crowded beats far
            if (!why) {
                for (int i = 0; i < n; i++) {
                    workers[i]->stats.steps_completed += 1;
                    workers[i]->stats.tile_frames += 1;
                }
                continue;
            }
            // some strip's tiles could not hold this frame: everybody still has its input (the other
            // buffer), packs it, and the frame is redone on k_phys / k_rebin below
            for (int i = 0; i < n; i++) {
                wrach_cuda_worker *w = on(i);
                w->tcur ^= 1;
                w->stats.tile_fallbacks++;
                if (3u - why == kTileWhyFar) w->tiles_retry_at = w->stats.steps_completed + 8;
                else w->tiles_off_until_upload = true;
                w->h_ctrl->tile_fail = 0;
                CU(cudaMemsetAsync(&w->ctrl->tile_fail, 0, 2 * sizeof(uint32_t), w->stream));
            }
        }
        // ---- the packed layout: physics, exchange of the leavers, re-bin
        Frame frames[64];
        for (int i = 0; i < n; i++) {
            wrach_cuda_worker *w = on(i);
            int rc = make_packed(w);
            if (rc) return report(w, rc);
            w->tiled_valid = false;
            frames[i] = make_frame(w, w->cur_enqueue);
            if (w->neighbour_mode) launch_nb_halo_pack(w, frames[i]);
        }
        if (workers[0]->neighbour_mode) {
            int rc = sync_all();
            if (rc) return rc;
        }
        for (int i = 0; i < n; i++) {
            wrach_cuda_worker *w = on(i);
            if (w->neighbour_mode) {
                int rc = nb_halo_exchange_peers(w);
                if (rc) return report(w, rc);
                launch_neighbours(w, frames[i]);
            }
            launch_phys(w, frames[i]);
            w->cur_enqueue ^= 1;
            w->pending += 1;
            CU(cudaMemcpyAsync(w->h_ctrl, w->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, w->stream));
        }
        int rc = sync_all();
        if (rc) return rc;
        bool far = false;
        for (int i = 0; i < n; i++) far = far || workers[i]->h_ctrl->far_seen || workers[i]->h_ctrl->abort;
        if (far) {
            // a far mover (or more leavers than a message holds) somewhere: nobody re-bins on the fast
            // path; the collective re-bin places every particle of every strip
            for (int i = 0; i < n; i++) {
                wrach_cuda_worker *w = on(i);
                const uint32_t completed = w->h_ctrl->steps_done - w->steps_done_seen;  // (earlier frames of this call)
                w->steps_done_seen = w->h_ctrl->steps_done;
                w->cur ^= (int)(completed & 1u);
                w->pending -= completed;
                w->stats.steps_completed += completed;
                if (w->h_ctrl->dense_seen) w->dense_enabled = true;
            }
            rc = strip_collective_rebin_group(workers, n);
            if (rc) return report(workers[0], rc);
            continue;
        }
        for (int i = 0; i < n; i++) {
            wrach_cuda_worker *w = on(i);
            for (int side = 0; side < 2; side++) {
                if (!w->peer[side]) continue;
                CU(cudaMemcpyAsync(w->imp_buf[side], w->peer[side]->exp_buf[side ^ 1], msg_bytes(w->exp_cap),
                                   cudaMemcpyDefault, w->stream));
                w->stats.halo_bytes_sent += msg_bytes(w->exp_cap);
            }
        }
        rc = sync_all();
        if (rc) return rc;
        for (int i = 0; i < n; i++) {
            wrach_cuda_worker *w = on(i);
            launch_rebin(w, frames[i]);
            CU(cudaMemcpyAsync(w->h_ctrl, w->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, w->stream));
            CU(cudaGetLastError());
        }
    }
    int first_rc = WRACH_OK;
    for (int i = 0; i < n; i++) {  // every strip is resolved, whatever happened to another; the first failure is reported
        wrach_cuda_worker *w = on(i);
        if (!w->pending && !w->tile_pending) continue;
        int rc = resolve(w);
        if (rc && !first_rc) first_rc = report(w, rc);
    }
    return first_rc;
}

// Strips are cut on tile boundaries (multiples of the tile width in cell columns) whenever the grid
// has at least two tile columns per strip, so that a strip's tiles are whole and its neighbours'
// edge tile columns can serve as its ghosts; narrower grids are split evenly by cell column and run
// on k_phys / k_rebin with the particle exchange.
void wrach_cuda_strip_columns(uint32_t grid_x, int rank, int n_ranks, uint32_t *begin, uint32_t *end) {
    if (n_ranks < 1) n_ranks = 1;
    const uint64_t tw = TileShape::TW, tile_cols = (grid_x + tw - 1) / tw;
    uint64_t b, e;
    if (tile_cols >= 2ull * (uint64_t)n_ranks) {
        b = std::min<uint64_t>(grid_x, tile_cols * (uint64_t)rank / (uint64_t)n_ranks * tw);
        e = std::min<uint64_t>(grid_x, tile_cols * (uint64_t)(rank + 1) / (uint64_t)n_ranks * tw);
    } else {
        b = (uint64_t)grid_x * (uint64_t)rank / (uint64_t)n_ranks;
        e = (uint64_t)grid_x * (uint64_t)(rank + 1) / (uint64_t)n_ranks;
    }
    if (begin) *begin = (uint32_t)b;
    if (end) *end = (uint32_t)e;
}

void wrach_cuda_destroy(wrach_cuda_worker *w) {
    if (!w) return;
    DeviceGuard guard(w->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    for (int i = 0; i < 2; i++) cudaFree(w->idx[i]);
    free_particle_buffer(w, 0, w->pos_in); free_particle_buffer(w, 1, w->vel_in); cudaFree(w->pos_out); cudaFree(w->vel_out);
    cudaFree(w->meta); cudaFree(w->cls); cudaFree(w->cls9); cudaFree(w->goff9); cudaFree(w->dense_list); cudaFree(w->run_total); cudaFree(w->run_base); cudaFree(w->vl_slot); cudaFree(w->vl_meta); cudaFree(w->vl_cnt); cudaFree(w->ctrl); cudaFree(w->tile_status);
    cudaFree(w->slow_cursor); cudaFree(w->slow_src); cudaFree(w->slow_ticket);
    for (int i = 0; i < 2; i++) {
        cudaFree(w->exp_buf[i]);
        cudaFree(w->imp_buf[i]);
    }
    cudaFree(w->imp_cnt);
    cudaFree(w->imp_off);
    for (int i = 0; i < 2; i++) {
        cudaFree(w->nb_send[i]);
        cudaFree(w->nb_recv[i]);
    }
    xr_free(w);
    cudaFree(w->xr_cnt);
    cudaFree(w->tile_vote);
    if (w->xr_host) cudaFreeHost(w->xr_host);
    if (w->h_tile_vote) cudaFreeHost(w->h_tile_vote);
    for (int i = 0; i < 2; i++) {
        cudaFree(w->tdata[i]);
        cudaFree(w->tstarts[i]);
    }
    if (w->comm_stream) {
        cudaStreamSynchronize(w->comm_stream);
        cudaStreamDestroy(w->comm_stream);
    }
    cudaFree(w->d_sig);
    if (w->ev_edge) cudaEventDestroy(w->ev_edge);
    if (w->ev_exch) cudaEventDestroy(w->ev_exch);
    if (w->comm && nccl_api() && nccl_api()->CommDestroy) nccl_api()->CommDestroy(w->comm);
    if (w->h_ctrl) cudaFreeHost(w->h_ctrl);
    for (auto e : w->ev)
        if (e) cudaEventDestroy(e);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
}

int wrach_cuda_write_slice(wrach_cuda_worker *w, wrach_buffer buffer, const void *src, size_t bytes) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    if (!src && bytes) return fail(w, WRACH_ERR_BAD_ARG, "null source");
    if (buffer == WRACH_WORLD_SETTINGS_UNIFORM) return fail(w, WRACH_ERR_BAD_ARG, "use wrach_cuda_write_settings");
    {  // uploads apply to the resolved state, in the packed layout
        int rc = settle_packed(w);
        if (rc) return rc;
    }
    size_t cap = 0;
    void *dst = buffer_ptr(w, buffer, &cap);
    if (!dst) return fail(w, WRACH_ERR_BAD_ARG, "unknown buffer %d", (int)buffer);
    if (bytes > cap) return fail(w, WRACH_ERR_CAPACITY, "write of %zu bytes into a %zu-byte buffer", bytes, cap);
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, w->stream));
    w->tiled_valid = false;             // the tiles are rebuilt from what was just written
    w->tiles_off_until_upload = false;  // new data: the tiles get another chance
    w->tiles_retry_at = 0;
    return WRACH_OK;
}

int wrach_cuda_write_settings(wrach_cuda_worker *w, const wrach_world_settings *settings) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    if (!settings) return fail(w, WRACH_ERR_BAD_ARG, "null settings");
    wrach_world_settings local = *settings;
    if (w->strip) {  // strips take the GLOBAL grid (as at creation) and the LOCAL particle count
        if (settings->grid_dimensions[0] != w->gs.grid_dimensions[0] || settings->grid_dimensions[1] != w->gs.grid_dimensions[1])
            return fail(w, WRACH_ERR_BAD_ARG, "strip workers cannot change the global grid");
        const uint32_t gcells = settings->grid_dimensions[0] * settings->grid_dimensions[1];
        int rcg = validate_settings(w, *settings, gcells + 2, 0xFFFFFFFFu);
        if (rcg) return rcg;
        w->gs = *settings;
        local.grid_dimensions[0] = w->col1 - w->col0;
        if (local.particles_in_frame_count > w->capacity)
            return fail(w, WRACH_ERR_CAPACITY, "particles_in_frame_count %u > capacity %u", local.particles_in_frame_count, w->capacity);
    } else {
        int rc = validate_settings(w, *settings, w->total_cells, w->capacity);
        if (rc) return rc;
    }
    if (memcmp(&w->s, &local, sizeof local) != 0) {  // (an unchanged uniform, as `tick` re-sends it, costs nothing)
        int rc = settle_packed(w);
        if (rc) return rc;
        w->tiled_valid = false;
    }
    w->s = local;  // the uniform travels by value with every kernel launch
    return WRACH_OK;
}

int wrach_cuda_step(wrach_cuda_worker *w, uint32_t n_steps) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    return enqueue_frames(w, n_steps, false, nullptr, nullptr);
}

int wrach_cuda_ready(wrach_cuda_worker *w) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    cudaError_t q = cudaStreamQuery(w->stream);
    if (q == cudaErrorNotReady) return 0;
    if (q != cudaSuccess) return fail(w, WRACH_ERR_CUDA, "stream error: %s", cudaGetErrorString(q));
    if (w->pending || w->tile_pending) {  // drained: account for the frames (and finish an aborted one if need be)
        int rc = resolve(w);
        if (rc) return rc;
        if (cudaStreamQuery(w->stream) == cudaErrorNotReady) return 0;  // recovery work was enqueued
    }
    return 1;
}

int wrach_cuda_sync(wrach_cuda_worker *w) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    return resolve(w);
}

static int read_common(wrach_cuda_worker *w, wrach_buffer buffer, void *dst, size_t bytes, bool wait) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    if (!dst && bytes) return fail(w, WRACH_ERR_BAD_ARG, "null destination");
    if (w->pending || w->tile_pending) {  // frames in flight: wait for them and learn how they went (one synchronisation)
        int rc = resolve(w);
        if (rc) return rc;
    }
    if (buffer != WRACH_WORLD_SETTINGS_UNIFORM) {  // buffers are read in the reference's packed layout
        int rc = make_packed(w);
        if (rc) return rc;
    }
    if (buffer == WRACH_WORLD_SETTINGS_UNIFORM) {
        if (bytes > sizeof(w->s)) return fail(w, WRACH_ERR_CAPACITY, "uniform is 32 bytes");
        memcpy(dst, &w->s, bytes);
        return WRACH_OK;
    }
    size_t cap = 0;
    void *src = buffer_ptr(w, buffer, &cap);
    if (!src) return fail(w, WRACH_ERR_BAD_ARG, "unknown buffer %d", (int)buffer);
    if (bytes > cap) return fail(w, WRACH_ERR_CAPACITY, "read of %zu bytes from a %zu-byte buffer", bytes, cap);
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, w->stream));
    if (wait) CU(cudaStreamSynchronize(w->stream));
    return WRACH_OK;
}

int wrach_cuda_read(wrach_cuda_worker *w, wrach_buffer buffer, void *dst, size_t bytes) {
    return read_common(w, buffer, dst, bytes, true);
}

int wrach_cuda_read_async(wrach_cuda_worker *w, wrach_buffer buffer, void *dst, size_t bytes) {
    return read_common(w, buffer, dst, bytes, false);
}

size_t wrach_cuda_buffer_bytes(const wrach_cuda_worker *w, wrach_buffer buffer) {
    if (!w) return 0;
    if (buffer == WRACH_WORLD_SETTINGS_UNIFORM) return sizeof(wrach_world_settings);
    size_t cap = 0;
    buffer_ptr(const_cast<wrach_cuda_worker *>(w), buffer, &cap);
    return cap;
}

void *wrach_cuda_device_pointer(wrach_cuda_worker *w, wrach_buffer buffer) {
    if (!w) return nullptr;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    if (settle_packed(w)) return nullptr;
    w->tiled_valid = false;  // the caller may write through the pointer
    size_t cap = 0;
    return buffer_ptr(w, buffer, &cap);
}

int wrach_cuda_settle(wrach_cuda_worker *w) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    int rc = settle_packed(w);
    if (rc) return rc;
    CU(cudaStreamSynchronize(w->stream));  // (the conversion kernels, if any ran)
    return WRACH_OK;
}

int wrach_cuda_export_buffer_fd(wrach_cuda_worker *w, wrach_buffer buffer, int *fd, size_t *alloc_bytes) {
    if (!w || !fd) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    if (buffer != WRACH_POSITIONS_IN && buffer != WRACH_VELOCITIES_IN)
        return fail(w, WRACH_ERR_BAD_ARG, "only POSITIONS_IN and VELOCITIES_IN can be exported (what a renderer binds, bind_groups.rs:61-83)");
    VmmApi *vm = vmm_api();
    if (!vm) return fail(w, WRACH_ERR_CUDA, "this driver does not offer the virtual memory management entry points");
    const int which = buffer == WRACH_POSITIONS_IN ? 0 : 1;
    wrach_cuda_worker::Shared &sh = w->shared[which];
    float2 **slot = which == 0 ? &w->pos_in : &w->vel_in;
    if (!sh.va) {
        // move the buffer, once, into a shareable allocation: the kernels take their pointers from the
        // worker at every launch, so nothing else changes
        int rc = settle_packed(w);
        if (rc) return rc;
        const CUmemAllocationProp prop = shareable_prop(w->device);
        size_t gran = 0;
        if (vm->GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
            return fail(w, WRACH_ERR_CUDA, "cuMemGetAllocationGranularity failed");
        const size_t want = ((size_t)w->capacity + 4) * sizeof(float2);  // (the padding the kernels' bulk copies rely on)
        const size_t bytes = (want + gran - 1) / gran * gran;
        CUmemGenericAllocationHandle handle = 0;
        if (vm->Create(&handle, bytes, &prop, 0) != CUDA_SUCCESS)
            return fail(w, WRACH_ERR_CUDA, "cuMemCreate of %zu shareable bytes failed", bytes);
        const CUdeviceptr va = vmm_map(vm, handle, bytes, w->device);
        if (!va) {
            vm->Release(handle);
            return fail(w, WRACH_ERR_CUDA, "mapping the shareable allocation failed");
        }
        CU(cudaMemsetAsync(reinterpret_cast<void *>(va), 0, bytes, w->stream));
        CU(cudaMemcpyAsync(reinterpret_cast<void *>(va), *slot, want, cudaMemcpyDeviceToDevice, w->stream));
        CU(cudaStreamSynchronize(w->stream));
        cudaFree(*slot);
        *slot = reinterpret_cast<float2 *>(va);
        sh.handle = handle;
        sh.va = va;
        sh.bytes = bytes;
    }
    int out = -1;
    if (vm->Export(&out, sh.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS || out < 0)
        return fail(w, WRACH_ERR_CUDA, "cuMemExportToShareableHandle failed");
    *fd = out;
    if (alloc_bytes) *alloc_bytes = sh.bytes;
    return WRACH_OK;
}

int wrach_cuda_selftest_import_fd(int device, int fd, size_t alloc_bytes, void *dst, size_t bytes) {
    wrach_cuda_worker *w = nullptr;  // errors go to the library-level message
    if (fd < 0 || !dst || bytes > alloc_bytes) return fail(nullptr, WRACH_ERR_BAD_ARG, "bad argument");
    DeviceGuard g(device);
    cudaFree(nullptr);  // (a context, if this is the process's first CUDA call)
    VmmApi *vm = vmm_api();
    if (!vm) return fail(nullptr, WRACH_ERR_CUDA, "this driver does not offer the virtual memory management entry points");
    CUmemGenericAllocationHandle handle = 0;
    const CUresult r = vm->Import(&handle, reinterpret_cast<void *>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (r != CUDA_SUCCESS) return fail(nullptr, WRACH_ERR_CUDA, "cuMemImportFromShareableHandle failed (%d)", (int)r);
    const CUdeviceptr va = vmm_map(vm, handle, alloc_bytes, device);
    if (!va) {
        vm->Release(handle);
        return fail(nullptr, WRACH_ERR_CUDA, "mapping the imported allocation failed");
    }
    const cudaError_t e = cudaMemcpy(dst, reinterpret_cast<void *>(va), bytes, cudaMemcpyDeviceToHost);
    vm->Unmap(va, alloc_bytes);
    vm->AddressFree(va, alloc_bytes);
    vm->Release(handle);
    if (e != cudaSuccess) return fail(w, WRACH_ERR_CUDA, "copy from the imported mapping failed: %s", cudaGetErrorString(e));
    return WRACH_OK;
}

const char *wrach_cuda_last_error(const wrach_cuda_worker *w) { return w ? w->err.c_str() : g_create_error.c_str(); }

void *wrach_cuda_alloc_host(size_t bytes) {
    void *p = nullptr;
    return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr;
}

void wrach_cuda_free_host(void *p) {
    if (p) cudaFreeHost(p);
}

int wrach_cuda_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return WRACH_ERR_BAD_ARG;
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? WRACH_OK : (cudaGetLastError(), WRACH_ERR_CUDA);
}

int wrach_cuda_host_unregister(void *p) {
    if (!p) return WRACH_ERR_BAD_ARG;
    return cudaHostUnregister(p) == cudaSuccess ? WRACH_OK : (cudaGetLastError(), WRACH_ERR_CUDA);
}

int wrach_cuda_set_neighbour_mode(wrach_cuda_worker *w, int enabled) {
    if (!w) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    int rc = resolve(w);  // frames already enqueued keep the mode they were enqueued with
    if (rc) return rc;
    if (enabled && w->strip && w->edge_mask && !w->nb_recv[0] && !w->nb_recv[1]) {
        // strips: the neighbours' edge columns arrive as ghost columns every frame (k_nb_halo_pack)
        const size_t bytes = nb_halo_bytes(w->s.grid_dimensions[1]);
        for (int i = 0; i < 2; i++) {
            if (!((w->edge_mask >> i) & 1u)) continue;
            CU(cudaMalloc(&w->nb_send[i], bytes));
            CU(cudaMalloc(&w->nb_recv[i], bytes));
            CU(cudaMemset(w->nb_send[i], 0, bytes));
            CU(cudaMemset(w->nb_recv[i], 0, bytes));
        }
    }
    w->neighbour_mode = enabled != 0;
    return WRACH_OK;
}

int wrach_cuda_step_timed(wrach_cuda_worker *w, uint32_t n_steps, float *elapsed_ms) {
    if (!w || !elapsed_ms) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    int rc = resolve(w);
    if (rc) return rc;
    const uint64_t slow_before = w->stats.slow_path_steps + w->stats.tile_fallbacks;
    CU(cudaEventRecord(w->ev[0], w->stream));
    rc = enqueue_frames(w, n_steps, false, nullptr, nullptr);
    if (rc) return rc;
    CU(cudaEventRecord(w->ev[1], w->stream));
    rc = resolve(w);
    if (rc) return rc;
    if (w->stats.slow_path_steps + w->stats.tile_fallbacks != slow_before) {  // recovery work ran after ev[1]: time up to now
        CU(cudaEventRecord(w->ev[1], w->stream));
    }
    CU(cudaEventSynchronize(w->ev[1]));
    CU(cudaEventElapsedTime(elapsed_ms, w->ev[0], w->ev[1]));
    return WRACH_OK;
}

int wrach_cuda_step_profiled(wrach_cuda_worker *w, uint32_t n_steps, float *phys_ms, float *rebin_ms) {
    if (!w || !phys_ms || !rebin_ms) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    int rc = resolve(w);
    if (rc) return rc;
    *phys_ms = *rebin_ms = 0.0f;
    rc = enqueue_frames(w, n_steps, true, phys_ms, rebin_ms);
    if (rc) return rc;
    rc = resolve(w);
    w->stats.last_phys_ms = n_steps ? *phys_ms / n_steps : 0.0f;
    w->stats.last_rebin_ms = n_steps ? *rebin_ms / n_steps : 0.0f;
    w->stats.phys_launches_last = w->stats.rebin_launches_last = n_steps;
    return rc;
}

#ifdef WRACH_TIMELINE
// debug builds only: run ONE frame with the phase timeline on and copy it out (16 stamps per block)
int wrach_cuda_debug_timeline(wrach_cuda_worker *w, unsigned long long *out, uint32_t n_blocks) {
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    int rc = resolve(w);
    if (rc) return rc;
    unsigned long long *buf = nullptr;
    CU(cudaMalloc(&buf, (size_t)n_blocks * 16 * sizeof(unsigned long long)));
    CU(cudaMemset(buf, 0, (size_t)n_blocks * 16 * sizeof(unsigned long long)));
    CU(cudaMemcpyToSymbol(wrach::g_timeline, &buf, sizeof(buf)));
    rc = enqueue_frames(w, 1, false, nullptr, nullptr);
    if (!rc) rc = resolve(w);
    CU(cudaMemcpy(out, buf, (size_t)n_blocks * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long *null = nullptr;
    CU(cudaMemcpyToSymbol(wrach::g_timeline, &null, sizeof(null)));
    cudaFree(buf);
    return rc;
}
#endif

#ifdef WRACH_DEBUG_PHYS_ONLY
// debug builds only (tools/ablate.py): launch k_phys alone `n` times on the current state and time it.
// The state is left as it was (k_phys only writes the *_out side), except the run totals.
int wrach_cuda_debug_phys_only(wrach_cuda_worker *w, uint32_t n, float *ms) {
    std::lock_guard<std::mutex> lock(w->mu);
    DeviceGuard g(w->device);
    int rc = resolve(w);
    if (rc) return rc;
    Frame f = make_frame(w, w->cur);
    for (int i = 0; i < 5; i++) launch_phys(w, f);
    CU(cudaEventRecord(w->ev[1], w->stream));
    for (uint32_t i = 0; i < n; i++) launch_phys(w, f);
    CU(cudaEventRecord(w->ev[2], w->stream));
    CU(cudaEventSynchronize(w->ev[2]));
    CU(cudaEventElapsedTime(ms, w->ev[1], w->ev[2]));
    *ms /= (float)n;
    CU(cudaMemsetAsync(w->run_total, 0, ((size_t)(w->cells + kRun - 1) / kRun + 1) * sizeof(uint32_t), w->stream));
    CU(cudaMemsetAsync(&w->ctrl->abort, 0, 2 * sizeof(uint32_t), w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return WRACH_OK;
}
#endif

int wrach_cuda_selftest_push_division(int device, unsigned long long *mismatches) {
    if (!mismatches) return WRACH_ERR_BAD_ARG;
    wrach_cuda_worker *w = nullptr;  // errors go to the library-level message
    DeviceGuard g(device);
    unsigned long long *d_bad = nullptr;
    CU(cudaMalloc(&d_bad, sizeof(*d_bad)));
    CU(cudaMemset(d_bad, 0, sizeof(*d_bad)));
    k_selftest_push_division<<<148 * 8, 256>>>(d_bad);
    CU(cudaMemcpy(mismatches, d_bad, sizeof(*d_bad), cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    return WRACH_OK;
}

int wrach_cuda_selftest_push_sqrt(int device, unsigned long long *mismatches) {
    if (!mismatches) return WRACH_ERR_BAD_ARG;
    wrach_cuda_worker *w = nullptr;  // errors go to the library-level message
    DeviceGuard g(device);
    unsigned long long *d_bad = nullptr;
    CU(cudaMalloc(&d_bad, sizeof(*d_bad)));
    CU(cudaMemset(d_bad, 0, sizeof(*d_bad)));
    k_selftest_push_sqrt<<<148 * 8, 256>>>(d_bad);
    CU(cudaMemcpy(mismatches, d_bad, sizeof(*d_bad), cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    return WRACH_OK;
}

int wrach_cuda_get_stats(wrach_cuda_worker *w, wrach_cuda_stats *out) {
    if (!w || !out) return WRACH_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(w->mu);
    *out = w->stats;
    return WRACH_OK;
}

}  // extern "C"
