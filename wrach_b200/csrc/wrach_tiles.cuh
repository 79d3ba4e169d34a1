// wrach_tiles.cuh — the fused frame: K1..K4 of the reference in ONE launch over 2-D tiles of cells.
//
// The reference's frame is physics -> count -> scan -> pack (runners/bevy/src/compute/builder.rs:86-89),
// with a global dependency between physics and pack: a particle's packed slot needs the prefix sum
// over ALL cells.  That dependency only exists because the packed layout is row-major over the whole
// grid.  Between two read-backs nobody sees the layout, so the worker keeps the state TILE-MAJOR:
// the grid is cut into tiles of TW x TH cells, every tile owns a fixed region of `tcap` slots
// (float4 = position + velocity per particle, cells row-major inside the tile) plus a table of
// u16 cell starts.  A particle moves at most one cell per frame (true once |v| <= 1), so everything
// that can land in a tile comes from the tile itself and the one-cell ring around it:
//
//   k_tile_frame  block = one tile.  Stages the tile (one TMA bulk copy) and its halo ring (vector
//                 loads from the eight neighbouring tiles) in shared memory, runs the reference's
//                 physics on ALL staged cells -- shaders/physics/src/cell.rs:52-95, particles.rs:62-107,
//                 particle.rs:46-92; one thread per cell over occupancy-sorted cells, exactly the
//                 arithmetic of k_phys -- then re-bins in shared memory: per source cell the size of
//                 each of its nine move classes and per particle its rank inside its class (running
//                 counters: the thread owns the cell), per destination cell the nine arrival groups
//                 in source-cell order (= ascending source slot of the reference's packed layout:
//                 the stable counting sort that is the canonical order of SURVEY.md section 8c), a
//                 block scan for the new cell starts, and one 16-byte store per particle into the
//                 tile's region of the other buffer.  The halo's physics is recomputed by every tile
//                 that borders it (+22 % cells for 30 x 14 tiles); in exchange a frame is one launch
//                 that reads and writes every particle once (32 N bytes instead of 64 N), with no
//                 global scan, no atomics on particle data and no inter-block dependency.
//   k_tile_unpack packed (reference layout) -> tiles, after an upload.
//   k_tile_pack_counts / k_slow_scan / k_tile_pack_copy   tiles -> packed, before a read-back.
//
// Anything the tiles cannot hold -- a particle moving further than one cell (first frames with
// |v| > cell size), a tile or its staged ring over capacity, a cell above 255 particles -- raises a
// sticky flag with the frame's ordinal: that frame's output is discarded (its input buffer is intact,
// frames are double-buffered), later frames are no-ops, and the host packs the input of the failed
// frame and replays from there with k_phys / k_rebin (wrach_worker.cu: resolve()).
#pragma once
#include "wrach_kernels.cuh"

namespace wrach {

constexpr uint32_t kTileWhyCrowded = 1;  // over capacity somewhere: tiles stay off until the next upload
constexpr uint32_t kTileWhyFar = 2;      // a far mover: tiles are retried a few frames later

struct TileFrame {
    Limits lim;
    uint32_t gx, gy;       // grid (cells)
    uint32_t ntx, nty;     // tiles
    uint32_t tcap;         // slots of one tile region
    uint32_t tss;          // u16 entries per tile in the starts tables (>= cells of a tile + 1, multiple of 8)
    uint32_t ord;          // ordinal of this frame: a failure stores ord + 1 in ctrl->tile_fail
    uint32_t pdl;          // let the next launch's blocks in early
    // Strip workers: the tile grid covers the strip's own cell columns plus one ghost tile column on
    // every side that has a neighbouring strip; tiles are numbered column by column (a tile column is
    // then one contiguous range of every array: what the ghost exchange sends), a launch covers the
    // tile columns [tx_first, tx_first + gridDim.x / nty), and tile-grid cell column 0 is global cell
    // column col0.  Single device: col_major = 0 (row by row), tx_first = 0, col0 = 0.
    uint32_t col_major, tx_first;
    uint32_t n_first, tx_second;  // a launch over up to three ranges of tile columns: n_first columns from tx_first,
    uint32_t n_second, tx_third;  //   n_second from tx_second, the rest from tx_third
    int32_t col0;
    // Strip workers over NCCL: the first n_edge_blocks blocks of the launch are the tile columns next
    // to a neighbouring strip.  They wait (spinning on a word the exchange stream writes) until the
    // ghosts of the input buffer have arrived, and count themselves done on another word the exchange
    // stream waits for -- so one launch per frame serves both, with nothing but kernels on the main stream.
    uint32_t n_edge_blocks, ghost_target;
    const uint32_t *ghost_ready;
    uint32_t *edge_done;
    const float2 *in_pos, *in_vel;  // [ntiles][tcap] each
    float2 *out_pos, *out_vel;
    const uint16_t *ts_in;  // [ntiles][tss]: [c] = first slot of local cell c inside the region, [NC] = particles in the tile
    uint16_t *ts_out;
    Ctrl *ctrl;
};

// linear index of tile (tx, ty): row by row on a single device, column by column on strip workers
__device__ __forceinline__ uint32_t tile_index(uint32_t col_major, uint32_t ntx, uint32_t nty, uint32_t tx, uint32_t ty) {
    return col_major ? tx * nty + ty : ty * ntx + tx;
}
__device__ __forceinline__ void tile_of_block(uint32_t col_major, uint32_t ntx, uint32_t nty, uint32_t tx_first, uint32_t b,
                                              uint32_t &tx, uint32_t &ty) {
    if (col_major) {
        tx = tx_first + b / nty;
        ty = b - (b / nty) * nty;
    } else {
        ty = b / ntx;
        tx = b - ty * ntx;
    }
}

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int TW, int TH>
struct TileGeo {
    static constexpr int EW = TW + 2, EH = TH + 2, EXT = EW * EH, NC = TW * TH, NH = EXT - NC;
    // halo cells in a fixed order: bottom row, top row, left column, right column
    __host__ __device__ static constexpr uint32_t halo_to_ext(uint32_t h) {
        return h < (uint32_t)EW ? h
             : h < 2u * EW ? (uint32_t)(EH - 1) * EW + (h - EW)
             : h < 2u * EW + TH ? (h - 2u * EW + 1u) * EW
             : (h - 2u * EW - TH + 1u) * EW + (EW - 1);
    }
    __host__ __device__ static constexpr uint32_t ext_to_halo(uint32_t ex, uint32_t ey) {
        return ey == 0 ? ex : ey == (uint32_t)EH - 1 ? EW + ex : ex == 0 ? 2u * EW + ey - 1u : 2u * EW + TH + ey - 1u;
    }
};

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float2 a, float2 b) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}

#ifndef WRACH_TILE_L2_AHEAD
#define WRACH_TILE_L2_AHEAD 148  // blocks ahead whose tile is prefetched into the L2 (0 = off)
#endif
#ifndef WRACH_TILE_RANK_SELP
#define WRACH_TILE_RANK_SELP 2
#endif
#ifndef WRACH_TILE_LATE_FAILCHECK
#define WRACH_TILE_LATE_FAILCHECK 1  // the "has an earlier frame failed" question travels with the table loads (-2 % on single-wave worlds)
#endif
#ifndef WRACH_TILE_GATHER_BATCH
#define WRACH_TILE_GATHER_BATCH 1  // ring gathers: all of a thread's loads issued before its first store (-2.9 % of the frame)
#endif
#ifndef WRACH_TILE_EAGER_TMA
#define WRACH_TILE_EAGER_TMA 0   // 1: bulk-copy the whole tile region at once instead of waiting for its population
#endif
template <int TW, int TH, int PCAP>
struct TileSmem {
    using G = TileGeo<TW, TH>;
    float2 pos[PCAP];                // staged particles: the tile's region verbatim, then the halo cells
    float2 vel[PCAP];                //   (two arrays: an 8-byte stride halves the bank conflicts of the per-cell walks)
    uint32_t meta[PCAP];             // per staged particle: destination cell << 16 | move code << 8 | rank; ~0 = leaves the tile
    uint32_t cnt9[G::EXT * 3];       // per staged cell: sizes of its nine move classes (bytes 0..8 of 12)
    uint32_t goff[G::NC * 3];        // per destination cell: first slot of each arrival group inside the cell (bytes 0..8)
    uint32_t hsrc[G::NH];            // halo cell: its first slot in the input buffer
    uint16_t hcnt[G::NH], hoff[G::NH];
    uint16_t est[G::EXT], en[G::EXT];  // staged cell: first particle in P, count
    uint16_t krank[G::EXT], order[G::EXT];
    uint16_t newstart[G::NC + 2];
    uint32_t bin[16], wsum[32];
    uint32_t n_own, n_halo, failed;
    __align__(8) uint64_t mbar;
};

// STRIP = false compiles the strip workers' extras out (column-major tile order, column ranges, the
// edge blocks' waits and counts, the global column offset): measured 1.4 % on the single-device frame.
template <int ARITH, int TW, int TH, int NT, int PCAP, int MINB, bool STRIP>
__global__ void __launch_bounds__(NT, MINB) k_tile_frame(const TileFrame tf) {
    using G = TileGeo<TW, TH>;
    constexpr uint32_t EW = G::EW, EXT = G::EXT, NC = G::NC, NH = G::NH;
    static_assert(EXT < 4096 && NT % 32 == 0 && NH <= 32 * 8, "tile shape");
    extern __shared__ __align__(128) uint8_t tile_smem_raw[];
    using TileS_ = TileSmem<TW, TH, PCAP>;
    TileS_ &sm = *reinterpret_cast<TileS_ *>(tile_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;

    // shared memory only above the wait: this block may be resident while the previous frame drains
    if (tid == 0) mbar_init(&sm.mbar, 1);
    if (tid < 16) sm.bin[tid] = 0;
    for (uint32_t i = tid; i < EXT * 3u; i += NT) sm.cnt9[i] = 0;
    __syncthreads();  // (the bins are added to right below)
    pdl_wait();
    if (tf.pdl) pdl_trigger();
    const bool edge_block = STRIP && blockIdx.x < tf.n_edge_blocks;
    const uint32_t col_major = STRIP ? tf.col_major : 0u;
    const int32_t col0 = STRIP ? tf.col0 : 0;
#if WRACH_TILE_LATE_FAILCHECK
    // Has an earlier frame (or the unpack) failed?  Thread 0 asks, and the block acts on the answer behind
    // the first barrier below: the question travels with the table loads instead of in front of them
    // (one round trip less at the start of every block; nothing before that barrier writes global memory).
    uint32_t failed0 = 0;
    if (tid == 0 || edge_block) failed0 = *(volatile uint32_t *)&tf.ctrl->tile_fail;
    if (edge_block && failed0) {  // (edge blocks answer at once: they must not wait for ghosts that never come)
        if (tid == 0) atomicAdd(tf.edge_done, 1u);  // (the exchange stream counts on every edge block)
        return;
    }
#else
    if (*(volatile uint32_t *)&tf.ctrl->tile_fail) {  // block-uniform: an earlier frame (or the unpack) failed
        if (edge_block && tid == 0) atomicAdd(tf.edge_done, 1u);  // (the exchange stream counts on every edge block)
        return;
    }
#endif
    if (edge_block) {  // the ghosts this block is about to read: has the exchange of the previous frame delivered them?
        if (tid == 0)
            while ((int32_t)(ld_acquire_sys_u32(tf.ghost_ready) - tf.ghost_target) < 0) __nanosleep(200);
        __syncthreads();
    }

    uint32_t tx, ty;
    tile_of_block(col_major, tf.ntx, tf.nty, tf.tx_first, blockIdx.x, tx, ty);
    if (col_major) {
        const uint32_t c = tx - tf.tx_first;
        if (c >= tf.n_first + tf.n_second) tx = tf.tx_third + (c - tf.n_first - tf.n_second);
        else if (c >= tf.n_first) tx = tf.tx_second + (c - tf.n_first);
    }
    const uint32_t T = tile_index(col_major, tf.ntx, tf.nty, tx, ty);
    const uint16_t *ts = tf.ts_in + (size_t)T * tf.tss;
#if WRACH_TILE_L2_AHEAD
    // The tile a block one SM-load further down the launch will stage: ask the L2 for its region and
    // start table now, so that block's two dependent round trips (tables, then particles) end in the
    // L2 instead of the HBM (-1.2 % of the frame at any distance from 64 to 222 blocks; the DRAM is 20 % busy).
    if (tid == 32) {
        const uint32_t b2 = blockIdx.x + (uint32_t)WRACH_TILE_L2_AHEAD;
        if (b2 < gridDim.x) {
            uint32_t tx2, ty2;
            tile_of_block(col_major, tf.ntx, tf.nty, tf.tx_first, b2, tx2, ty2);
            if (col_major) {
                const uint32_t c = tx2 - tf.tx_first;
                if (c >= tf.n_first + tf.n_second) tx2 = tf.tx_third + (c - tf.n_first - tf.n_second);
                else if (c >= tf.n_first) tx2 = tf.tx_second + (c - tf.n_first);
            }
            const uint32_t T2 = tile_index(col_major, tf.ntx, tf.nty, tx2, ty2);
            l2_prefetch(tf.in_pos + (size_t)T2 * tf.tcap, tf.tcap * 8u);
            l2_prefetch(tf.in_vel + (size_t)T2 * tf.tcap, tf.tcap * 8u);
            l2_prefetch(tf.ts_in + (size_t)T2 * tf.tss, (tf.tss * 2u) & ~15u);
        }
    }
#endif
    if (tid == 0) {
#if WRACH_TILE_EAGER_TMA
        // the whole region, whatever it holds: the copy starts now, not one round trip from now (the
        // tile's population arrives with the tables below); the ring is staged behind the region
        const uint32_t bytes = tf.tcap * 8u;
        mbar_expect_tx(&sm.mbar, 2u * bytes);
        tma_load_1d(sm.pos, tf.in_pos + (size_t)T * tf.tcap, bytes, &sm.mbar);
        tma_load_1d(sm.vel, tf.in_vel + (size_t)T * tf.tcap, bytes, &sm.mbar);
        sm.n_own = ts[NC];
#else
#if WRACH_TILE_LATE_FAILCHECK
        sm.failed = failed0;
        const uint32_t n_own = failed0 ? 0u : ts[NC];  // (no bulk copy behind a block that is about to leave)
#else
        const uint32_t n_own = ts[NC];
#endif
        sm.n_own = n_own;
        if (n_own) {  // (regions start on 16-byte boundaries; an odd count copies one slot of padding)
            const uint32_t bytes = ((n_own + 1u) & ~1u) * 8u;
            mbar_expect_tx(&sm.mbar, 2u * bytes);
            tma_load_1d(sm.pos, tf.in_pos + (size_t)T * tf.tcap, bytes, &sm.mbar);
            tma_load_1d(sm.vel, tf.in_vel + (size_t)T * tf.tcap, bytes, &sm.mbar);
        }
#endif
    }
    // ---- table of staged cells: ext cell (ex, ey) is grid cell (x0 + ex, y0 + ey); the tile's own
    // cells are ex in [1, TW], ey in [1, TH], the ring around them comes from the neighbouring tiles
    const int32_t x0 = (int32_t)(tx * TW) - 1, y0 = (int32_t)(ty * TH) - 1;
    for (uint32_t e = tid; e < EXT; e += NT) {
        const uint32_t ey = e / EW, ex = e - ey * EW;
        uint32_t n = 0;
        if (ex - 1u < (uint32_t)TW && ey - 1u < (uint32_t)TH) {
            const uint32_t lc = (ey - 1u) * TW + (ex - 1u), s0 = ts[lc];
            n = ts[lc + 1] - s0;
            sm.est[e] = (uint16_t)s0;
        } else {
            const uint32_t h = G::ext_to_halo(ex, ey);
            const uint32_t cx = (uint32_t)(x0 + (int32_t)ex), cy = (uint32_t)(y0 + (int32_t)ey);  // below zero wraps and fails the test
            uint32_t src = 0;
            if (cx < tf.gx && cy < tf.gy) {
                const uint32_t ntx_ = cx / TW, nty_ = cy / TH, t2 = tile_index(col_major, tf.ntx, tf.nty, ntx_, nty_);
                const uint32_t lc = (cy - nty_ * TH) * TW + (cx - ntx_ * TW);
                const uint16_t *t2s = tf.ts_in + (size_t)t2 * tf.tss;
                const uint32_t s0 = t2s[lc];
                n = t2s[lc + 1] - s0;
                // (a neighbouring strip whose tile overflowed sends a table that points past the region: that
                // frame has failed over there and everything computed from it is discarded -- just stay in bounds)
                if (STRIP && s0 + n > tf.tcap) n = 0;
                src = t2 * tf.tcap + s0;  // (regions total below 2^32 slots: checked by the host)
            }
            sm.hcnt[h] = (uint16_t)n;
            sm.hsrc[h] = src;
        }
        sm.en[e] = (uint16_t)n;
        const uint32_t key = 15u - min(n, 15u);  // fullest first; key 15 = empty
        sm.krank[e] = (uint16_t)((key << 12) | atomicAdd(&sm.bin[key], 1u));
    }
    __syncthreads();
#if WRACH_TILE_LATE_FAILCHECK
    if (sm.failed) return;  // block-uniform; thread 0 issued no bulk copy
#endif
    if (wid == 0) {  // where each halo cell's particles go in P, behind the tile's own
        constexpr uint32_t PER = (NH + 31u) / 32u;
        uint32_t v[PER], s = 0;
#pragma unroll
        for (uint32_t q = 0; q < PER; q++) {
            const uint32_t h = lane * PER + q;
            v[q] = h < NH ? sm.hcnt[h] : 0u;
            s += v[q];
        }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        uint32_t off = inc - s;
#pragma unroll
        for (uint32_t q = 0; q < PER; q++) {
            const uint32_t h = lane * PER + q;
            if (h < NH) sm.hoff[h] = (uint16_t)min(off, 0xFFFFu);
            off += v[q];
        }
        if (lane == 31) sm.n_halo = inc;
    }
    // cells sorted by occupancy (counting sort on min(count, 15)): a warp's 32 cells then need about
    // the same number of pair slots and particle trips -- pushes are serial per cell, one thread each
    for (uint32_t e = tid; e < EXT; e += NT) {
        const uint32_t kr = sm.krank[e], key = kr >> 12;
        uint32_t before = 0;
#pragma unroll
        for (uint32_t q = 0; q < 15; q++) before += q < key ? sm.bin[q] : 0u;
        sm.order[before + (kr & 4095u)] = (uint16_t)e;
    }
    __syncthreads();
#if WRACH_TILE_EAGER_TMA
    const uint32_t n_own = tf.tcap, n_ext = n_own + sm.n_halo, n_mine = sm.n_own;  // ring behind the region
    constexpr bool kBulk = true;
#else
    // (an odd tile population was copied with one slot of padding: the ring starts behind it)
    const uint32_t n_own = (sm.n_own + 1u) & ~1u, n_ext = n_own + sm.n_halo, n_mine = sm.n_own;
    const bool kBulk = n_own != 0u;
#endif
    uint32_t why = 0;
    if (n_ext > (uint32_t)PCAP) {  // the ring does not fit the stage (block-uniform)
        if (tid == 0) {
            if (kBulk) mbar_wait(&sm.mbar, 0);  // never leave a bulk copy in flight behind us
            tf.ctrl->tile_why = kTileWhyCrowded;
            tf.ctrl->tile_fail = tf.ord + 1u;
            if (edge_block) atomicAdd(tf.edge_done, 1u);
        }
        return;
    }
#if !WRACH_TILE_EAGER_TMA
    if (tid == 0 && (sm.n_own & 1u)) sm.meta[sm.n_own] = 0xFFFFFFFFu;  // the padding slot holds no particle
#endif
    for (uint32_t h = tid; h < NH; h += NT) sm.est[G::halo_to_ext(h)] = (uint16_t)(n_own + sm.hoff[h]);
#if WRACH_TILE_GATHER_BATCH
    {  // sixteen threads per halo cell, one particle each: all of a thread's loads are issued before its first store
        constexpr uint32_t R = (NH * 16u + NT - 1u) / NT;
        float2 gp[R], gv[R];
        uint32_t gd[R];
#pragma unroll
        for (uint32_t q = 0; q < R; q++) {
            const uint32_t i = tid + q * NT, h = i >> 4, k = i & 15u;
            gd[q] = 0xFFFFFFFFu;
            if (i < NH * 16u && k < sm.hcnt[h]) {
                const uint32_t src = sm.hsrc[h] + k;
                gd[q] = n_own + sm.hoff[h] + k;
                gp[q] = __ldg(tf.in_pos + src);
                gv[q] = __ldg(tf.in_vel + src);
            }
        }
#pragma unroll
        for (uint32_t q = 0; q < R; q++) {
            if (gd[q] != 0xFFFFFFFFu) {
                sm.pos[gd[q]] = gp[q];
                sm.vel[gd[q]] = gv[q];
            }
        }
        for (uint32_t i = tid; i < NH * 16u; i += NT) {  // (a halo cell with more than sixteen particles: the rest)
            const uint32_t h = i >> 4, n = sm.hcnt[h];
            if (n > 16u) {
                const uint32_t dst = n_own + sm.hoff[h], src = sm.hsrc[h];
                for (uint32_t k = (i & 15u) + 16u; k < n; k += 16u) {
                    sm.pos[dst + k] = __ldg(tf.in_pos + src + k);
                    sm.vel[dst + k] = __ldg(tf.in_vel + src + k);
                }
            }
        }
    }
#else
    for (uint32_t i = tid; i < NH * 16u; i += NT) {  // sixteen threads per halo cell, one particle each (and again beyond 16)
        const uint32_t h = i >> 4, n = sm.hcnt[h];
        uint32_t k = i & 15u;
        if (k < n) {
            const uint32_t dst = n_own + sm.hoff[h], src = sm.hsrc[h];
            for (; k < n; k += 16u) {
                sm.pos[dst + k] = __ldg(tf.in_pos + src + k);
                sm.vel[dst + k] = __ldg(tf.in_vel + src + k);
            }
        }
    }
#endif
    if (kBulk) mbar_wait(&sm.mbar, 0);
    __syncthreads();

    // ---- physics + classification: one thread per staged cell, in place (see k_phys: the same loop).
    // Warps take groups of 32 cells of the sorted order in a snake (w, 2 NW - 1 - w, 2 NW + w, ...) so
    // that a warp that got full cells first gets emptier ones next.
    const Limits &L = tf.lim;
    const uint32_t n_live = EXT - sm.bin[15], n_groups = (n_live + 31u) / 32u;
    constexpr uint32_t NW = NT / 32;
    bool far = false, crowded = false;
    for (uint32_t pass = 0;; pass++) {
        const uint32_t g = pass * NW + ((pass & 1u) ? NW - 1u - wid : wid);
        if (g >= n_groups) break;
        const uint32_t idx = g * 32u + lane;
        if (idx >= n_live) continue;
        const uint32_t e = sm.order[idx];
        const uint32_t ey = e / EW, ex = e - ey * EW;
        const uint32_t n = sm.en[e], n9 = min(n, (uint32_t)kMaxInCell);
        if (n > 255u) crowded = true;  // 8-bit ranks and class sizes
        const CellBox box = make_cell_box(L, __fmul_rn((float)(col0 + x0 + (int32_t)ex), L.cs), __fmul_rn((float)(y0 + (int32_t)ey), L.cs));  // exact
        // which of the nine moves end inside the tile, and the local index of the cell one step down-left
        const uint32_t mx = (ex >= 2u ? 1u : 0u) | (ex - 1u < (uint32_t)TW ? 2u : 0u) | (ex + 1u <= (uint32_t)TW ? 4u : 0u);
        const uint32_t my = (ey >= 2u ? 1u : 0u) | (ey - 1u < (uint32_t)TH ? 2u : 0u) | (ey + 1u <= (uint32_t)TH ? 4u : 0u);
        const uint32_t mask9 = ((my & 1u) ? mx : 0u) | ((my & 2u) ? mx << 3 : 0u) | ((my & 4u) ? mx << 6 : 0u);
        const int32_t dbase = ((int32_t)ey - 2) * TW + (int32_t)ex - 2;
        const uint32_t s0 = sm.est[e];
        constexpr int kVelOff = (int)(offsetof(TileS_, vel) - offsetof(TileS_, pos));
        uint32_t Pi = smem_u32(sm.pos) + s0 * 8u;  // position of the cell's particle i (its velocity sits kVelOff behind)
        uint32_t Mi = smem_u32(sm.meta) + s0 * 4u;
        // running sizes of the nine move classes, a byte each: codes 0-3, 4-7 and 8 (registers: the
        // shared-memory pipe is what this loop is short of)
        uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < n; i++, Pi += 8u, Mi += 4u) {
            float2 pi = lds_f2<0>(Pi);
            if (i + 1 < n9) {
                const uint32_t partners = n9 - i;  // u = 1 .. partners - 1
                // two registers take turns as "this partner" / "the next one": the next partner is fetched
                // before this one is pushed (no push touches it), with no register moves between the slots
                float2 pa = lds_f2<8>(Pi), pb = pa;
#define WRACH_TILE_PAIR_SLOT(U, CUR, NXT)                                           \
    {                                                                               \
        if ((uint32_t)(U) + 1u < partners) NXT = lds_f2<8 * ((U) + 1)>(Pi);         \
        if (push_pair<ARITH>(pi, CUR)) sts_f2<8 * (U)>(Pi, CUR);                    \
        if ((uint32_t)(U) + 1u >= partners) goto row_done;                          \
    }
                WRACH_TILE_PAIR_SLOT(1, pa, pb) WRACH_TILE_PAIR_SLOT(2, pb, pa) WRACH_TILE_PAIR_SLOT(3, pa, pb)
                WRACH_TILE_PAIR_SLOT(4, pb, pa) WRACH_TILE_PAIR_SLOT(5, pa, pb) WRACH_TILE_PAIR_SLOT(6, pb, pa)
                WRACH_TILE_PAIR_SLOT(7, pa, pb)
                if (push_pair<ARITH>(pi, pb)) sts_f2<64>(Pi, pb);  // u = 8: the last partner of row 0 of a full cell
#undef WRACH_TILE_PAIR_SLOT
            row_done:;
            }
            const float2 v0 = lds_f2<kVelOff>(Pi);
            float2 v = v0;
            uint32_t ddx1, ddy1;
            const uint32_t code = finish_in_box(L, box, pi, v, ddx1, ddy1);
            if (code > 8u) far = true;
            // rank inside the (cell, move) class = the class's running size (a far mover, code 15, counts nowhere)
#if WRACH_TILE_RANK_SELP == 2
            // codes 0..7 as the eight bytes of one 64-bit register pair (c0 | c1 << 32), code 8 in c2: PTX
            // shifts by 64 or more give zero, so codes 8 and 15 add nothing to and read nothing from the pair
            uint32_t rank;
            {
                const uint32_t sh8 = code * 8u;
                unsigned long long c64 = ((unsigned long long)c1 << 32) | c0, inc64, r64;
                asm("shl.b64 %0, %1, %2;" : "=l"(inc64) : "l"(1ull), "r"(sh8));
                asm("shr.b64 %0, %1, %2;" : "=l"(r64) : "l"(c64), "r"(sh8));
                c64 += inc64;
                c0 = (uint32_t)c64;
                c1 = (uint32_t)(c64 >> 32);
                const bool is8 = code == 8u;
                rank = is8 ? c2 : ((uint32_t)r64 & 255u);
                c2 += is8 ? 1u : 0u;
            }
#else
            const uint32_t sh = (code & 3u) * 8u, hi = code >> 2, inc = 1u << sh;
#if WRACH_TILE_RANK_SELP
            uint32_t cc;  // (two selects; left to itself the compiler branches here, with a reconvergence region around it: -2 % of the frame)
            asm("{\n.reg .pred p, q;\nsetp.eq.u32 p, %1, 0;\nsetp.eq.u32 q, %1, 1;\nselp.u32 %0, %3, %4, q;\nselp.u32 %0, %2, %0, p;\n}"
                : "=&r"(cc) : "r"(hi), "r"(c0), "r"(c1), "r"(c2));
            const uint32_t rank = (cc >> sh) & 255u;
#else
            const uint32_t rank = ((hi == 0u ? c0 : hi == 1u ? c1 : c2) >> sh) & 255u;
#endif
            c0 += hi == 0u ? inc : 0u;
            c1 += hi == 1u ? inc : 0u;
            c2 += hi == 2u ? inc : 0u;
#endif
            uint32_t m = 0xFFFFFFFFu;
            if (code <= 8u && ((mask9 >> code) & 1u)) m = ((uint32_t)(dbase + (int32_t)(ddy1 * TW + ddx1)) << 16) | (code << 8) | rank;
            sts_f2<0>(Pi, pi);
            // a velocity only changes at the world's edge (a sign flip) or while it is above the speed limit
            if (__float_as_uint(v.x) != __float_as_uint(v0.x) || __float_as_uint(v.y) != __float_as_uint(v0.y)) sts_f2<kVelOff>(Pi, v);
            sts_u32(Mi, m);
        }
        const uint32_t Ci = smem_u32(sm.cnt9) + e * 12u;
        sts_u32(Ci, c0);
        sts_u32(Ci + 4u, c1);
        sts_u32(Ci + 8u, c2);
    }
    if (far) why = kTileWhyFar;
    __syncthreads();

    // ---- destination cells: nine arrival groups in source-cell order (group g comes from the cell at
    // (-ddx, -ddy) with move code 8 - g), sizes from the class counters
    constexpr uint32_t K = (NC + NT - 1) / NT;
    uint32_t sz[K], sum = 0;
#pragma unroll
    for (uint32_t q = 0; q < K; q++) {
        const uint32_t d = tid * K + q;
        sz[q] = 0;
        if (d < NC) {
            const uint32_t ly = d / TW, lx = d - ly * TW;
            const uint8_t *c9 = reinterpret_cast<const uint8_t *>(sm.cnt9) + ((ly + 1u) * EW + lx + 1u) * 12u;
            uint32_t acc = 0, w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
            for (int g = 0; g < 9; g++) {
                const int code = 8 - g, ddy = code / 3 - 1, ddx = code % 3 - 1;
                const uint32_t c = c9[(-(ddy * (int)EW) - ddx) * 12 + code];
                const uint32_t a = acc & 255u;
                if (g < 4) w0 |= a << (8 * g);
                else if (g < 8) w1 |= a << (8 * (g - 4));
                else w2 = a;
                acc += c;
            }
            if (acc > 255u) crowded = true;
            sm.goff[d * 3u] = w0;
            sm.goff[d * 3u + 1u] = w1;
            sm.goff[d * 3u + 2u] = w2;
            sz[q] = acc;
        }
        sum += sz[q];
    }
    uint32_t total;
    uint32_t base = block_exclusive_scan<NT>(sum, sm.wsum, total);
    uint16_t *ts_out = tf.ts_out + (size_t)T * tf.tss;
#pragma unroll
    for (uint32_t q = 0; q < K; q++) {
        const uint32_t d = tid * K + q;
        if (d < NC) {
            sm.newstart[d] = (uint16_t)min(base, 0xFFFFu);
            ts_out[d] = (uint16_t)min(base, 0xFFFFu);
            base += sz[q];
        }
    }
    if (tid == 0) ts_out[NC] = (uint16_t)min(total, 0xFFFFu);
    if (total > tf.tcap) {
        crowded = true;
    }
    if (crowded) why = kTileWhyCrowded;
    if (why) {
        tf.ctrl->tile_why = why;
        tf.ctrl->tile_fail = tf.ord + 1u;
    }
    if (total > tf.tcap) {  // block-uniform: the stores below would leave the region
        if (edge_block && tid == 0) atomicAdd(tf.edge_done, 1u);
        return;
    }
    __syncthreads();

    // ---- every staged particle that ends in the tile goes to its slot
    float2 *out_pos = tf.out_pos + (size_t)T * tf.tcap, *out_vel = tf.out_vel + (size_t)T * tf.tcap;
    const uint8_t *goff8 = reinterpret_cast<const uint8_t *>(sm.goff);
    // (the tile's own particles, then the ring -- which starts at n_own, past the padding / the region)
    // (B particles per thread and trip, their shared-memory reads batched: +1 % for B = 2, +2 % for B = 4)
    for (uint32_t i0 = tid; i0 < n_mine + (n_ext - n_own); i0 += NT) {
        const uint32_t i = i0 < n_mine ? i0 : n_own + (i0 - n_mine);
        const uint32_t m = sm.meta[i];
        if (m != 0xFFFFFFFFu) {
            const uint32_t d = m >> 16, g = 8u - ((m >> 8) & 15u);
            const uint32_t slot = sm.newstart[d] + goff8[d * 12u + g] + (m & 255u);
            out_pos[slot] = sm.pos[i];
            out_vel[slot] = sm.vel[i];
        }
    }
    if (edge_block) {  // this tile's output is what a neighbouring strip is waiting for
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(tf.edge_done, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// packed (reference layout: indices / positions_in / velocities_in) -> tiles

struct TileConv {
    uint32_t gx, gy, ntx, nty, tcap, tss, ord, cells;  // gx, gy: the PACKED grid (a strip's own columns)
    uint32_t col_major, tx_first;  // see TileFrame; blocks cover tile columns from tx_first on
    uint32_t x_off;                // tile-grid cell column of packed column 0 (a strip's left ghost column, if any)
    uint32_t capacity;             // slots of the packed particle buffers
    uint32_t *idx;          // packed `indices` (reference layout: [k + 1] = first slot of cell k)
    float2 *pos, *vel;      // packed positions_in / velocities_in
    float2 *tpos, *tvel;    // [ntiles][tcap] each
    uint16_t *ts;
    Ctrl *ctrl;
};

template <int TW, int TH>
__global__ void __launch_bounds__(256) k_tile_unpack(const TileConv c) {
    constexpr uint32_t NC = TW * TH, NT = 256, K = (NC + NT - 1) / NT;
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t row_src[TH], row_dst[TH + 1];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t tx, ty;
    tile_of_block(c.col_major, c.ntx, c.nty, c.tx_first, blockIdx.x, tx, ty);
    const uint32_t T = tile_index(c.col_major, c.ntx, c.nty, tx, ty);
    uint32_t cnt[K], sum = 0;
    bool crowded = false;
#pragma unroll
    for (uint32_t q = 0; q < K; q++) {
        const uint32_t lc = tid * K + q;
        cnt[q] = 0;
        if (lc < NC) {
            const uint32_t ly = lc / TW, lx = lc - ly * TW, x = tx * TW + lx - c.x_off, y = ty * TH + ly;  // (x wraps below the strip)
            if (x < c.gx && y < c.gy) {
                const uint32_t k = y * c.gx + x, s0 = c.idx[k + 1];
                cnt[q] = c.idx[k + 2] - s0;
                if (lx == 0) row_src[ly] = s0;
            } else if (lx == 0) {
                row_src[ly] = 0;
            }
            if (cnt[q] > 255u) crowded = true;
        }
        sum += cnt[q];
    }
    uint32_t total;
    uint32_t base = block_exclusive_scan<NT>(sum, wsum, total);
    uint16_t *ts = c.ts + (size_t)T * c.tss;
#pragma unroll
    for (uint32_t q = 0; q < K; q++) {
        const uint32_t lc = tid * K + q;
        if (lc < NC) {
            ts[lc] = (uint16_t)min(base, 0xFFFFu);
            if (lc % TW == 0) row_dst[lc / TW] = base;
            base += cnt[q];
        }
    }
    if (tid == 0) {
        ts[NC] = (uint16_t)min(total, 0xFFFFu);
        row_dst[TH] = total;
    }
    if (total > c.tcap) crowded = true;
    if (crowded) {
        c.ctrl->tile_why = kTileWhyCrowded;
        c.ctrl->tile_fail = c.ord + 1u;
    }
    if (total > c.tcap) return;  // block-uniform
    __syncthreads();
    // a row of the tile is one contiguous range of the packed arrays and of the region
    float2 *dpos = c.tpos + (size_t)T * c.tcap, *dvel = c.tvel + (size_t)T * c.tcap;
    for (uint32_t ly = wid; ly < (uint32_t)TH; ly += NT / 32) {
        const uint32_t s = row_src[ly], d0 = row_dst[ly], n = row_dst[ly + 1] - d0;
        for (uint32_t j = lane; j < n; j += 32) {
            dpos[d0 + j] = c.pos[s + j];
            dvel[d0 + j] = c.vel[s + j];
        }
    }
}

// tiles -> packed, step 1: the size of every cell at [k + 2] (an inclusive scan then leaves the
// reference layout: [k + 1] = first slot of cell k, [C + 1] = N)
template <int TW, int TH>
__global__ void __launch_bounds__(256) k_tile_pack_counts(const TileConv c) {
    constexpr uint32_t NC = TW * TH;
    uint32_t tx, ty;
    tile_of_block(c.col_major, c.ntx, c.nty, c.tx_first, blockIdx.x, tx, ty);
    const uint32_t T = tile_index(c.col_major, c.ntx, c.nty, tx, ty);
    const uint16_t *ts = c.ts + (size_t)T * c.tss;
    if (blockIdx.x == 0 && threadIdx.x < 2) c.idx[threadIdx.x] = 0;
    for (uint32_t lc = threadIdx.x; lc < NC; lc += blockDim.x) {
        const uint32_t ly = lc / TW, lx = lc - ly * TW, x = tx * TW + lx - c.x_off, y = ty * TH + ly;
        if (x < c.gx && y < c.gy) c.idx[y * c.gx + x + 2] = (uint32_t)ts[lc + 1] - (uint32_t)ts[lc];
    }
}

// step 3 (after the scan): copy every tile row to its place
template <int TW, int TH>
__global__ void __launch_bounds__(256) k_tile_pack_copy(const TileConv c) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t tx, ty;
    tile_of_block(c.col_major, c.ntx, c.nty, c.tx_first, blockIdx.x, tx, ty);
    const uint32_t T = tile_index(c.col_major, c.ntx, c.nty, tx, ty);
    const uint16_t *ts = c.ts + (size_t)T * c.tss;
    // a strip's population changes from frame to frame: it must still fit the packed buffers
    if (c.idx[c.cells + 1] > c.capacity) {
        if (blockIdx.x == 0 && tid == 0) c.ctrl->strip_error = 2u;
        return;
    }
    const float2 *spos = c.tpos + (size_t)T * c.tcap, *svel = c.tvel + (size_t)T * c.tcap;
    for (uint32_t ly = wid; ly < (uint32_t)TH; ly += blockDim.x / 32) {
        const uint32_t y = ty * TH + ly, x = tx * TW - c.x_off;
        if (y >= c.gy || x >= c.gx) continue;
        const uint32_t s0 = ts[ly * TW], n = (uint32_t)ts[ly * TW + TW] - s0, d0 = c.idx[y * c.gx + x + 1];
        for (uint32_t j = lane; j < n; j += 32) {
            c.pos[d0 + j] = spos[s0 + j];
            c.vel[d0 + j] = svel[s0 + j];
        }
    }
}

}  // namespace wrach
