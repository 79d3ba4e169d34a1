// wrach_xrebin.cuh — strips: the re-bin of ONE frame across all strips, for any displacement.
//
// The reference's K2..K4 (assets/shaders/particles_per_cell.wgsl:7-30, prefix_sum.wgsl:17-123,
// pack_new_particle_data.wgsl:10-45) place a particle wherever its new position keys to, however
// far it flew: velocities are clamped only AFTER the integration (shaders/physics/src/
// particles.rs:102-104), so a first frame with |v| above the cell size is legal.  On one device that
// is the generic re-bin of wrach_kernels.cuh (k_slow_*).  On strips the destination may belong to
// any other strip, so the frame's re-bin becomes a collective step:
//
//   k_xr_count   per source slot: owner of the destination column -> particles per destination strip
//   (host)       the count matrix of all strips (NCCL all-gather, or read directly when the strips
//                share a process) -> segment offsets, capacity check
//   k_xr_pack    records (position, velocity, 64-bit source key) grouped by destination strip
//   (exchange)   ncclSend/ncclRecv of the segments (device copies in-process)
//   k_xr_cells / k_slow_scan / k_xr_scatter / k_xr_place
//                counts per local cell, the new `indices`, and the canonical order inside a cell:
//                ascending source key = (global source cell, slot inside it) = ascending slot of the
//                single-device packed layout -- the stable counting sort every other path produces.
//
// A rare path (first frames after an upload with wild velocities, a burst of pushes): written for
// clarity, not for the roofline.
#pragma once
#include "wrach_kernels.cuh"

namespace wrach {

constexpr int kMaxStrips = 64;

struct XRec {                    // one particle on its way to the strip that owns its new cell
    float2 p, v;
    unsigned long long key;      // (global source cell << 32) | slot inside the source cell
};
static_assert(sizeof(XRec) == 24, "exchange record");

struct XRebin {
    wrach_world_settings gs;     // the GLOBAL grid and view rectangle
    uint32_t lgx, col0;          // this strip: columns, first global column
    uint32_t cells;              // lgx * grid.y
    uint32_t n_ranks, rank;
    uint32_t col_end[kMaxStrips];   // global column where strip r ends
    uint32_t send_off[kMaxStrips];  // first record of the segment for strip r in `send`
    const uint32_t *starts;      // indices the frame's physics read: source cell of every slot, and N
    const float2 *pos_out, *vel_out;
    uint32_t *send_cnt;          // [n_ranks]: particles per destination strip (k_xr_count)
    uint32_t *send_cursor;       // [n_ranks]: k_xr_pack's claim counters (zeroed)
    XRec *send, *recv;
    uint32_t n_recv;             // records this strip received (its new population)
    uint32_t *starts_next, *cursor, *src;
    unsigned long long *key_at;  // source key of the record k_xr_scatter put at each slot
    float2 *pos_in, *vel_in;
};

__device__ __forceinline__ uint32_t xr_global_col(const XRebin &x, float px) {
    return min(cell_coord(px, x.gs.view_anchor[0], (float)x.gs.cell_size), x.gs.grid_dimensions[0] - 1u);
}
__device__ __forceinline__ uint32_t xr_row(const XRebin &x, float py) {
    return min(cell_coord(py, x.gs.view_anchor[1], (float)x.gs.cell_size), x.gs.grid_dimensions[1] - 1u);
}
__device__ __forceinline__ uint32_t xr_owner(const XRebin &x, uint32_t col) {
    uint32_t r = 0;
    while (r + 1 < x.n_ranks && col >= x.col_end[r]) r++;
    return r;
}
// cell k of the local grid with starts[k+1] <= slot < starts[k+2] (empty cells share a start)
__device__ __forceinline__ uint32_t xr_source_cell(const XRebin &x, uint32_t slot) {
    uint32_t lo = 0, hi = x.cells;  // first k in [0, cells] with starts[k+1] > slot, minus one
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (x.starts[mid + 1] > slot) hi = mid;
        else lo = mid + 1;
    }
    return lo - 1u;
}

// One trip of a warp over 32 source slots: destination strip per lane, and per group of lanes with
// the same destination one atomicAdd on `counter[dest]`; returns the lane's place in its group's claim.
__device__ __forceinline__ uint32_t xr_claim(uint32_t *counter, bool active, uint32_t dest) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t peers = __match_any_sync(0xffffffffu, active ? dest : 0xFFFFFFFFu);
    const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
    uint32_t base = 0;
    if (active && lane == leader) base = atomicAdd(&counter[dest], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(peers & lanes_below(lane));
}

__global__ void __launch_bounds__(256) k_xr_count(const XRebin x) {
    const uint32_t n = x.starts[x.cells + 1];
    const uint32_t stride = gridDim.x * blockDim.x, rounds = (n + stride - 1) / stride;
    for (uint32_t t = 0; t < rounds; t++) {
        const uint32_t i = t * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool active = i < n;
        const uint32_t dest = active ? xr_owner(x, xr_global_col(x, x.pos_out[i].x)) : 0u;
        xr_claim(x.send_cnt, active, dest);
    }
}

__global__ void __launch_bounds__(256) k_xr_pack(const XRebin x) {
    const uint32_t n = x.starts[x.cells + 1];
    const uint32_t stride = gridDim.x * blockDim.x, rounds = (n + stride - 1) / stride;
    const uint32_t ggx = x.gs.grid_dimensions[0];
    for (uint32_t t = 0; t < rounds; t++) {
        const uint32_t i = t * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool active = i < n;
        float2 p = make_float2(0.f, 0.f);
        uint32_t dest = 0;
        if (active) {
            p = x.pos_out[i];
            dest = xr_owner(x, xr_global_col(x, p.x));
        }
        const uint32_t at = xr_claim(x.send_cursor, active, dest);
        if (active) {
            const uint32_t k = xr_source_cell(x, i);
            const uint32_t row = k / x.lgx, col = k - row * x.lgx;
            XRec r;
            r.p = p;
            r.v = x.vel_out[i];
            r.key = ((unsigned long long)(row * ggx + x.col0 + col) << 32) | (unsigned long long)(i - x.starts[k + 1]);
            x.send[x.send_off[dest] + at] = r;
        }
    }
}

__device__ __forceinline__ uint32_t xr_local_cell(const XRebin &x, float2 p) {
    return xr_row(x, p.y) * x.lgx + (xr_global_col(x, p.x) - x.col0);
}

// counts land at [cell + 2] so that the inclusive scan leaves [k+1] = first slot of cell k
__global__ void __launch_bounds__(256) k_xr_cells(const XRebin x) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < x.n_recv; j += gridDim.x * blockDim.x)
        atomicAdd(&x.starts_next[xr_local_cell(x, x.recv[j].p) + 2], 1u);
}

__global__ void __launch_bounds__(256) k_xr_scatter(const XRebin x) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < x.n_recv; j += gridDim.x * blockDim.x) {
        const XRec r = x.recv[j];
        const uint32_t c = xr_local_cell(x, r.p);
        const uint32_t slot = x.starts_next[c + 1] + atomicAdd(&x.cursor[c], 1u);
        x.src[slot] = j;
        x.key_at[slot] = r.key;
    }
}

// canonical order inside a cell: ascending source key (keys are unique)
__global__ void __launch_bounds__(256) k_xr_place(const XRebin x) {
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < x.n_recv; d += gridDim.x * blockDim.x) {
        const XRec r = x.recv[x.src[d]];
        const uint32_t c = xr_local_cell(x, r.p);
        const uint32_t b = x.starts_next[c + 1], e = x.starts_next[c + 2];
        uint32_t rank = 0;
        for (uint32_t q = b; q < e; q++) rank += x.key_at[q] < r.key;
        x.pos_in[b + rank] = r.p;
        x.vel_in[b + rank] = r.v;
    }
}

// ---------------------------------------------------------------------------------------------
// Tile frames on strips: what this strip puts into the vote that ends a batch of frames.  The
// reduction is a maximum, so the EARLIEST failed frame of any strip wins: vote[0] = ~(ordinal + 1),
// vote[1] = why (crowded beats far: the more conservative retry policy).
__global__ void k_tile_vote(const Ctrl *ctrl, uint32_t *vote) {
    const uint32_t failed = ctrl->tile_fail;
    vote[0] = failed ? ~failed : 0u;
    vote[1] = failed ? 3u - ctrl->tile_why : 0u;  // kTileWhyCrowded (1) -> 2, kTileWhyFar (2) -> 1
}

// ---------------------------------------------------------------------------------------------
// Ghost columns of the opt-in 3x3 neighbour mode (wrach_cuda_set_neighbour_mode on strip workers):
// the first-nine positions of every cell of the strip's two edge columns, as the neighbouring strip's
// k_neighbours needs them for its own edge column.  Layout of one side's message: grid.y counts
// (u32, padded to an even number), then grid.y x 9 positions.
__host__ __device__ inline size_t nb_halo_bytes(uint32_t gy) { return ((size_t)((gy + 1u) & ~1u)) * 4 + (size_t)gy * kMaxInCell * sizeof(float2); }
__device__ __forceinline__ const float2 *nb_halo_pos(const uint8_t *base, uint32_t gy) {
    return reinterpret_cast<const float2 *>(base + ((size_t)((gy + 1u) & ~1u)) * 4);
}

__global__ void __launch_bounds__(256) k_nb_halo_pack(const Frame f, uint8_t *out_left, uint8_t *out_right) {
    if (f.ctrl->abort) return;
    const uint32_t gx = f.s.grid_dimensions[0], gy = f.s.grid_dimensions[1];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < 2u * gy * kMaxInCell; e += gridDim.x * blockDim.x) {
        const uint32_t side = e / (gy * kMaxInCell), rest = e - side * gy * kMaxInCell;
        const uint32_t row = rest / kMaxInCell, k = rest - row * kMaxInCell;
        uint8_t *out = side == 0 ? out_left : out_right;
        if (!out) continue;
        const uint32_t c = row * gx + (side == 0 ? 0u : gx - 1u);
        const uint32_t s0 = f.starts[c + 1], n9 = min(f.starts[c + 2] - s0, (uint32_t)kMaxInCell);
        if (k == 0) reinterpret_cast<uint32_t *>(out)[row] = n9;
        if (k < n9) const_cast<float2 *>(nb_halo_pos(out, gy))[(size_t)row * kMaxInCell + k] = f.pos_in[s0 + k];
    }
}

}  // namespace wrach
