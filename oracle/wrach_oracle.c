/* wrach_oracle.c — see wrach_oracle.h.  TEST INFRASTRUCTURE ONLY (checker + CPU baseline).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off keeps every a*b+c unfused unless fmaf() is spelled out, which is how the two
 * arithmetic variants are told apart.  No -ffast-math anywhere.
 */
#include "wrach_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PREFIX_SUM_HACK 1u   /* shaders/physics/src/lib.rs:30 */
#define MIN_DISTANCE 1.0f    /* shaders/physics/src/particles.rs:18 */

/* ------------------------------------------------------------------------------------------ */
/* host-side twin                                                                              */

/* Rust `as i32` from f32: saturating, NaN -> 0. */
static int32_t f32_as_i32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}

/* std f32::div_euclid: q = trunc(a / b); if a % b < 0 { q - 1 (b > 0) | q + 1 }  */
static float f32_div_euclid(float a, float b) {
    float q = truncf(a / b);
    if (fmodf(a, b) < 0.0f) return b > 0.0f ? q - 1.0f : q + 1.0f;
    return q;
}

/* spatial_bin.rs:48-64 */
int32_t wo_cell_coord(float position, uint16_t cell_size) {
    return f32_as_i32(f32_div_euclid(position, (float)cell_size));
}

/* spatial_bin.rs:68-89: cells from get_cell_coord(viewport.xy) ..= get_cell_coord(viewport.zw) */
void wo_active_grid(const float viewport[4], uint16_t cell_size, int32_t bottom_left[2], uint32_t grid[2]) {
    int32_t blx = wo_cell_coord(viewport[0], cell_size), bly = wo_cell_coord(viewport[1], cell_size);
    int32_t trx = wo_cell_coord(viewport[2], cell_size), try_ = wo_cell_coord(viewport[3], cell_size);
    bottom_left[0] = blx;
    bottom_left[1] = bly;
    /* grid.y counts every y in the range; grid.x is counted on the last row only, so an empty
     * y-range leaves both at 0 (spatial_bin.rs:77-86). */
    grid[1] = try_ >= bly ? (uint32_t)(try_ - bly + 1) : 0u;
    grid[0] = (try_ >= bly && trx >= blx) ? (uint32_t)(trx - blx + 1) : 0u;
}

/* particle_store.rs:116-133 */
uint32_t wo_max_particles_per_frame(uint32_t total_cells, uint16_t cell_size) {
    uint32_t per_cell = (uint32_t)cell_size * (uint32_t)cell_size;
    uint32_t normally = total_cells * per_cell;
    uint32_t one_percent = (normally + 99u) / 100u;
    return normally + 10u * one_percent;
}

/* particle_store.rs:54-59 (bucket by div_euclid cell, insertion order inside a cell) followed by
 * spatial_bin.rs:103-149 (walk active cells row-major, two leading zeros, running totals). */
uint32_t wo_create_packed_data(const float viewport[4], uint16_t cell_size, const float *p, uint32_t n,
                               uint32_t *indices, float *positions, float *velocities) {
    int32_t bl[2];
    uint32_t grid[2];
    wo_active_grid(viewport, cell_size, bl, grid);
    uint32_t cells = grid[0] * grid[1];
    uint32_t *cursor = (uint32_t *)calloc((size_t)cells + 1, sizeof(uint32_t));
    int64_t *slot_cell = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
    for (uint32_t i = 0; i < n; i++) {
        int64_t cx = (int64_t)wo_cell_coord(p[4 * i + 0], cell_size) - bl[0];
        int64_t cy = (int64_t)wo_cell_coord(p[4 * i + 1], cell_size) - bl[1];
        if (cx < 0 || cy < 0 || cx >= (int64_t)grid[0] || cy >= (int64_t)grid[1]) {
            slot_cell[i] = -1; /* stays in the store, not in this frame */
            continue;
        }
        slot_cell[i] = cy * (int64_t)grid[0] + cx;
        cursor[slot_cell[i]]++;
    }
    indices[0] = 0;
    indices[1] = 0;
    uint32_t running = 0;
    for (uint32_t c = 0; c < cells; c++) {
        uint32_t cnt = cursor[c];
        cursor[c] = running;
        running += cnt;
        indices[c + 2] = running;
    }
    for (uint32_t i = 0; i < n; i++) {
        if (slot_cell[i] < 0) continue;
        uint32_t d = cursor[slot_cell[i]]++;
        positions[2 * d] = p[4 * i];
        positions[2 * d + 1] = p[4 * i + 1];
        velocities[2 * d] = p[4 * i + 2];
        velocities[2 * d + 1] = p[4 * i + 3];
    }
    free(cursor);
    free(slot_cell);
    return running;
}

/* ------------------------------------------------------------------------------------------ */
/* device-side restatement                                                                     */

static uint32_t f32_to_u32_sat(float v) {
    if (!(v > 0.0f)) return 0u; /* negatives and NaN */
    if (v >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)v;
}

/* particles_per_cell.wgsl:14-27 == pack_new_particle_data.wgsl:19-33 */
uint32_t wo_cell_key(const wo_settings *s, float x, float y) {
    float rx = x - s->view_anchor[0];
    float ry = y - s->view_anchor[1];
    float cs = (float)s->cell_size;
    uint32_t cx = f32_to_u32_sat(floorf(rx / cs));
    uint32_t cy = f32_to_u32_sat(floorf(ry / cs));
    return cy * s->grid_dimensions[0] + cx;
}

/* One pair of particles.rs:62-94 (+ glam Vec2::distance): L and R are pushed apart in place. */
static inline void push_close_pair(float *L, float *R, int arith) {
    float dx = L[0] - R[0], dy = L[1] - R[1]; /* left.distance(right) = (left-right).length() */
    float d2 = arith == WO_ARITH_SPV ? fmaf(dx, dx, dy * dy) : dx * dx + dy * dy;
    float distance = sqrtf(d2);
    if (distance > MIN_DISTANCE) return;      /* particles.rs:70-72 */
    if (distance == 0.0f) distance = 0.0001f; /* particles.rs:74-76 */
    /* particles.rs:85-94 */
    float force = 0.5f * (MIN_DISTANCE - distance) / distance;
    float vx = R[0] - L[0], vy = R[1] - L[1];
    if (arith == WO_ARITH_SPV) {
        float lx = fmaf(-vx, force, L[0]), ly = fmaf(-vy, force, L[1]);
        float rx = fmaf(vx, force, R[0]), ry = fmaf(vy, force, R[1]);
        L[0] = lx; L[1] = ly; R[0] = rx; R[1] = ry;
    } else {
        vx *= force;
        vy *= force;
        L[0] -= vx; L[1] -= vy;
        R[0] += vx; R[1] += vy;
    }
}

/* particles.rs:62-83.  p = x0,y0,x1,y1,... */
void wo_pairs(float *p, uint32_t count, int arith) {
    for (uint32_t l = 0; l < count; l++)
        for (uint32_t r = l + 1; r < count; r++) push_close_pair(p + 2 * l, p + 2 * r, arith);
}

/* particle.rs:80-82, 46-70, 73-77 in the order of particles.rs:102-104 */
static inline void integrate_and_limit(const wo_settings *s, float *px, float *py, float *vx, float *vy) {
    float x = *px + *vx, y = *py + *vy, ux = *vx, uy = *vy;
    float x0 = s->view_anchor[0], y0 = s->view_anchor[1];
    float x1 = s->view_anchor[0] + s->view_dimensions[0], y1 = s->view_anchor[1] + s->view_dimensions[1];
    if (x > x1) { x = x1; ux *= -1.0f; }
    if (x < x0) { x = x0; ux *= -1.0f; }
    if (y > y1) { y = y1; uy *= -1.0f; }
    if (y < y0) { y = y0; uy *= -1.0f; }
    /* f32::clamp(-1, 1): NaN stays NaN */
    if (ux < -1.0f) ux = -1.0f;
    if (ux > 1.0f) ux = 1.0f;
    if (uy < -1.0f) uy = -1.0f;
    if (uy > 1.0f) uy = 1.0f;
    *px = x; *py = y; *vx = ux; *vy = uy;
}

/* One cell of K1 minus the clear (cell.rs:52-95).  `start`,`count` already read. */
static inline void physics_for_cell(const wo_settings *s, uint32_t start, uint32_t all_count, const float *pos_in,
                                    const float *vel_in, float *pos_out, float *vel_out, int arith) {
    uint32_t count = all_count > WO_MAX_PARTICLES_IN_CELL ? WO_MAX_PARTICLES_IN_CELL : all_count;
    float p[2 * WO_MAX_PARTICLES_IN_CELL], v[2 * WO_MAX_PARTICLES_IN_CELL];
    for (uint32_t i = 0; i < count; i++) { /* particles.rs:48-56 */
        p[2 * i] = pos_in[2 * (size_t)(start + i)];
        p[2 * i + 1] = pos_in[2 * (size_t)(start + i) + 1];
        v[2 * i] = vel_in[2 * (size_t)(start + i)];
        v[2 * i + 1] = vel_in[2 * (size_t)(start + i) + 1];
    }
    wo_pairs(p, count, arith);
    for (uint32_t i = 0; i < count; i++) { /* particles.rs:96-107 */
        integrate_and_limit(s, &p[2 * i], &p[2 * i + 1], &v[2 * i], &v[2 * i + 1]);
        size_t d = start + i; /* written back to the SAME slot, particle.rs:85-92 */
        pos_out[2 * d] = p[2 * i]; pos_out[2 * d + 1] = p[2 * i + 1];
        vel_out[2 * d] = v[2 * i]; vel_out[2 * d + 1] = v[2 * i + 1];
    }
    for (uint32_t i = count; i < all_count; i++) { /* cell.rs:79-95: overflow, no collisions */
        size_t d = start + i;
        float x = pos_in[2 * d], y = pos_in[2 * d + 1], ux = vel_in[2 * d], uy = vel_in[2 * d + 1];
        integrate_and_limit(s, &x, &y, &ux, &uy);
        pos_out[2 * d] = x; pos_out[2 * d + 1] = y;
        vel_out[2 * d] = ux; vel_out[2 * d + 1] = uy;
    }
}

/* lib.rs:41-64 + cell.rs:52-76,99-131.  Work items run in ascending order: item c reads
 * indices[c+1], indices[c+2] and only then zeroes indices[c+1], so the reference's unordered
 * cross-workgroup clear (latent race) never bites. */
void wo_k1_physics(const wo_settings *s, uint32_t *indices, const float *pos_in, const float *vel_in,
                   float *pos_out, float *vel_out, int arith) {
    uint32_t cells = s->grid_dimensions[0] * s->grid_dimensions[1];
    uint32_t last_cell = cells + PREFIX_SUM_HACK - 1; /* cell.rs:53-55 */
    for (uint32_t id = 0; id < cells; id++) {
        uint32_t current = id + PREFIX_SUM_HACK;
        if (current > last_cell) break;
        uint32_t start = indices[current], marker = indices[current + 1];
        physics_for_cell(s, start, marker - start, pos_in, vel_in, pos_out, vel_out, arith);
        indices[current] = 0;                               /* cell.rs:117 */
        if (current == last_cell) indices[current + 1] = 0; /* cell.rs:120-121 */
    }
}

void wo_k2_count(const wo_settings *s, const float *pos_out, uint32_t *indices) {
    for (uint32_t i = 0; i < s->particles_in_frame_count; i++)
        indices[wo_cell_key(s, pos_out[2 * (size_t)i], pos_out[2 * (size_t)i + 1])] += 1u;
}

void wo_k3_scan(const wo_settings *s, uint32_t *indices) {
    uint32_t total = s->grid_dimensions[0] * s->grid_dimensions[1] + 2u; /* prefix_sum.wgsl:23 */
    uint32_t running = 0;
    for (uint32_t i = 0; i < total; i++) {
        uint32_t v = indices[i];
        indices[i] = running;
        running += v;
    }
}

void wo_k4_pack(const wo_settings *s, const float *pos_out, const float *vel_out, uint32_t *indices,
                float *pos_in, float *vel_in) {
    for (uint32_t k = s->particles_in_frame_count; k-- > 0;) { /* arrival order: descending index */
        uint32_t cell_index = wo_cell_key(s, pos_out[2 * (size_t)k], pos_out[2 * (size_t)k + 1]) + 1u;
        uint32_t count = indices[cell_index]; /* atomicSub returns the old value */
        indices[cell_index] = count - 1u;
        size_t d = count - 1u;
        pos_in[2 * d] = pos_out[2 * (size_t)k]; pos_in[2 * d + 1] = pos_out[2 * (size_t)k + 1];
        vel_in[2 * d] = vel_out[2 * (size_t)k]; vel_in[2 * d + 1] = vel_out[2 * (size_t)k + 1];
    }
}

void wo_step(const wo_settings *s, uint32_t *indices, float *pos_in, float *vel_in, float *pos_out,
             float *vel_out, uint32_t steps, int arith) {
    for (uint32_t t = 0; t < steps; t++) { /* builder.rs:86-89 */
        wo_k1_physics(s, indices, pos_in, vel_in, pos_out, vel_out, arith);
        wo_k2_count(s, pos_out, indices);
        wo_k3_scan(s, indices);
        wo_k4_pack(s, pos_out, vel_out, indices, pos_in, vel_in);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* EXTENSION, not in the reference: the 3x3 neighbour search
 * the reference's module comment announces ("the physics for a cell (and its surroundings)",
 * cell.rs:1-2) but never implements.  Opt-in; every reference-parity check runs with it off.
 *
 * Before K1, every particle i among the first nine of its cell (the reference's per-cell capacity,
 * cell.rs:21,29-30) meets every particle j among the first nine of each of the eight surrounding
 * cells -- cells in row-major order (dy = -1, 0, 1; dx = -1, 0, 1; the centre skipped), slots
 * ascending -- through push_close_particles_apart (particles.rs:62-94), with j at its frame-start
 * position and only i's half of the push kept.  i accumulates its pushes one after the other
 * (Gauss-Seidel in i, Jacobi in j), so no particle's result depends on another's: any execution
 * order gives the same bits.  Own-cell pairs stay K1's business, untouched.  pos_tmp is scratch. */
static void neighbour_cell(const wo_settings *s, const uint32_t *indices, const float *pos_in, float *pos_tmp,
                           uint32_t c, int arith) {
    const uint32_t gx = s->grid_dimensions[0], gy = s->grid_dimensions[1];
    const uint32_t start = indices[c + 1], all = indices[c + 2] - start;
    const uint32_t n9 = all > WO_MAX_PARTICLES_IN_CELL ? WO_MAX_PARTICLES_IN_CELL : all;
    const uint32_t cy = c / gx, cx = c - cy * gx;
    for (uint32_t k = 0; k < n9; k++) {
        float me[2] = {pos_in[2 * (size_t)(start + k)], pos_in[2 * (size_t)(start + k) + 1]};
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (dx == 0 && dy == 0) continue;
                const int64_t nx = (int64_t)cx + dx, ny = (int64_t)cy + dy;
                if (nx < 0 || ny < 0 || nx >= (int64_t)gx || ny >= (int64_t)gy) continue;
                const uint32_t nc = (uint32_t)ny * gx + (uint32_t)nx;
                const uint32_t ns = indices[nc + 1], nall = indices[nc + 2] - ns;
                const uint32_t m9 = nall > WO_MAX_PARTICLES_IN_CELL ? WO_MAX_PARTICLES_IN_CELL : nall;
                for (uint32_t j = 0; j < m9; j++) {
                    float other[2] = {pos_in[2 * (size_t)(ns + j)], pos_in[2 * (size_t)(ns + j) + 1]};
                    push_close_pair(me, other, arith); /* other's half is dropped */
                }
            }
        pos_tmp[2 * (size_t)(start + k)] = me[0];
        pos_tmp[2 * (size_t)(start + k) + 1] = me[1];
    }
}

void wo_neighbour_pass(const wo_settings *s, const uint32_t *indices, float *pos_in, float *pos_tmp, int arith,
                       int threads) {
    const uint32_t cells = s->grid_dimensions[0] * s->grid_dimensions[1];
    (void)threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
#endif
    for (int64_t c = 0; c < (int64_t)cells; c++) neighbour_cell(s, indices, pos_in, pos_tmp, (uint32_t)c, arith);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
#endif
    for (int64_t c = 0; c < (int64_t)cells; c++) { /* commit: only the first nine of a cell took part */
        const uint32_t start = indices[c + 1], all = indices[c + 2] - start;
        const uint32_t n9 = all > WO_MAX_PARTICLES_IN_CELL ? WO_MAX_PARTICLES_IN_CELL : all;
        for (uint32_t k = 0; k < n9; k++) {
            pos_in[2 * (size_t)(start + k)] = pos_tmp[2 * (size_t)(start + k)];
            pos_in[2 * (size_t)(start + k) + 1] = pos_tmp[2 * (size_t)(start + k) + 1];
        }
    }
}

/* `steps` frames of neighbour pass + K1..K4.  threads = 1: everything serial. */
int wo_step_neighbours(const wo_settings *s, uint32_t *indices, float *pos_in, float *vel_in, float *pos_out,
                       float *vel_out, uint32_t steps, int arith, int threads) {
    int used = 1;
    for (uint32_t t = 0; t < steps; t++) {
        wo_neighbour_pass(s, indices, pos_in, pos_out, arith, threads);
        if (threads == 1) wo_step(s, indices, pos_in, vel_in, pos_out, vel_out, 1, arith);
        else used = wo_step_parallel(s, indices, pos_in, vel_in, pos_out, vel_out, 1, arith, threads);
    }
    return used;
}

/* ------------------------------------------------------------------------------------------ */
/* OpenMP variant: the CPU baseline.  Same arithmetic, same canonical order.                    */

int wo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int wo_step_parallel(const wo_settings *s, uint32_t *indices, float *pos_in, float *vel_in, float *pos_out,
                     float *vel_out, uint32_t steps, int arith, int threads) {
#ifndef _OPENMP
    (void)threads;
    wo_step(s, indices, pos_in, vel_in, pos_out, vel_out, steps, arith);
    return 1;
#else
    if (threads <= 0) threads = omp_get_max_threads();
    const uint32_t cells = s->grid_dimensions[0] * s->grid_dimensions[1];
    const uint32_t n = s->particles_in_frame_count;
    uint32_t *keys = (uint32_t *)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
    uint32_t *count = (uint32_t *)malloc((size_t)(cells + 2) * sizeof(uint32_t));
    uint32_t *lo = (uint32_t *)malloc((size_t)threads * sizeof(uint32_t));
    uint32_t *span = (uint32_t *)malloc((size_t)threads * sizeof(uint32_t));
    uint32_t **win = (uint32_t **)calloc((size_t)threads, sizeof(uint32_t *));
    uint32_t *partial = (uint32_t *)malloc((size_t)(threads + 1) * sizeof(uint32_t));

    for (uint32_t t = 0; t < steps; t++) {
        /* K1: cells are independent (cell.rs:52-76 touches only its own slot range); the clear of
         * `indices` is subsumed by rebuilding the whole array below. */
#pragma omp parallel for schedule(static, 512) num_threads(threads)
        for (uint32_t c = 0; c < cells; c++) {
            uint32_t start = indices[c + 1], marker = indices[c + 2];
            physics_for_cell(s, start, marker - start, pos_in, vel_in, pos_out, vel_out, arith);
        }
        /* K2 (+ the rank half of K4) as a parallel STABLE counting sort: thread w owns the contiguous
         * chunk [w*n/T, (w+1)*n/T) and a private histogram over the key window that chunk touches. */
#pragma omp parallel num_threads(threads)
        {
            int w = omp_get_thread_num();
            uint32_t b = (uint32_t)((uint64_t)n * (uint64_t)w / (uint64_t)threads);
            uint32_t e = (uint32_t)((uint64_t)n * (uint64_t)(w + 1) / (uint64_t)threads);
            uint32_t kmin = UINT32_MAX, kmax = 0;
            for (uint32_t i = b; i < e; i++) {
                uint32_t k = wo_cell_key(s, pos_out[2 * (size_t)i], pos_out[2 * (size_t)i + 1]);
                keys[i] = k;
                if (k < kmin) kmin = k;
                if (k > kmax) kmax = k;
            }
            lo[w] = kmin;
            span[w] = b == e ? 0u : kmax - kmin + 1u;
            win[w] = span[w] ? (uint32_t *)calloc(span[w], sizeof(uint32_t)) : NULL;
            for (uint32_t i = b; i < e; i++) win[w][keys[i] - kmin]++;
#pragma omp for schedule(static)
            for (uint32_t k = 0; k < cells + 2; k++) count[k] = 0;
        }
        /* Threads in ascending order (= ascending source slot): count[] accumulates the cell totals
         * (K2's result) while each private entry becomes that thread's rank offset inside the cell. */
        for (int q = 0; q < threads; q++) {
            uint32_t kmin = lo[q], sp = span[q];
            uint32_t *wq = win[q];
#pragma omp parallel for schedule(static) num_threads(threads)
            for (uint32_t j = 0; j < sp; j++) {
                uint32_t c = wq[j], r = count[kmin + j];
                wq[j] = r;
                count[kmin + j] = r + c;
            }
        }
        /* K3: exclusive scan (prefix_sum.wgsl semantics), two-level. count[k] -> start of cell k. */
#pragma omp parallel num_threads(threads)
        {
            int w = omp_get_thread_num();
            uint32_t tot = cells + 2u;
            uint32_t sb = (uint32_t)((uint64_t)tot * (uint64_t)w / (uint64_t)threads);
            uint32_t se = (uint32_t)((uint64_t)tot * (uint64_t)(w + 1) / (uint64_t)threads);
            uint32_t sum = 0;
            for (uint32_t i = sb; i < se; i++) sum += count[i];
            partial[w + 1] = sum;
#pragma omp barrier
#pragma omp single
            {
                partial[0] = 0;
                for (int q = 0; q < threads; q++) partial[q + 1] += partial[q];
            }
            uint32_t running = partial[w];
            for (uint32_t i = sb; i < se; i++) {
                uint32_t v = count[i];
                count[i] = running;
                running += v;
            }
#pragma omp barrier
            /* K4: destination = start(cell) + rank; ranks ascend with the source slot. */
            uint32_t b = (uint32_t)((uint64_t)n * (uint64_t)w / (uint64_t)threads);
            uint32_t e = (uint32_t)((uint64_t)n * (uint64_t)(w + 1) / (uint64_t)threads);
            for (uint32_t i = b; i < e; i++) {
                size_t d = (size_t)count[keys[i]] + win[w][keys[i] - lo[w]]++;
                pos_in[2 * d] = pos_out[2 * (size_t)i]; pos_in[2 * d + 1] = pos_out[2 * (size_t)i + 1];
                vel_in[2 * d] = vel_out[2 * (size_t)i]; vel_in[2 * d + 1] = vel_out[2 * (size_t)i + 1];
            }
            free(win[w]);
            win[w] = NULL;
            /* indices after K4: [0] = 0, [k+1] = start of cell k, [C+1] = N */
#pragma omp for schedule(static)
            for (uint32_t k = 0; k <= cells; k++) indices[k + 1] = count[k];
        }
        indices[0] = 0;
    }
    free(keys); free(count); free(lo); free(span); free(win); free(partial);
    return threads;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* scene generator                                                                             */

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static float unit24(uint64_t seed, uint64_t id, uint64_t comp) {
    return (float)(splitmix64(seed ^ splitmix64(id * 4ull + comp)) >> 40) * (1.0f / 16777216.0f);
}

void wo_generate_scene(uint64_t seed, uint64_t first_id, uint32_t n, float width, float height, int pile,
                       float *out) {
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; i++) {
        uint64_t id = first_id + i;
        float ux = unit24(seed, id, 0), uy = unit24(seed, id, 1);
        if (pile) uy = (uy * uy) * (uy * uy);
        out[4 * (size_t)i + 0] = ux * width;
        out[4 * (size_t)i + 1] = uy * height;
        out[4 * (size_t)i + 2] = unit24(seed, id, 2) - 0.5f;
        out[4 * (size_t)i + 3] = unit24(seed, id, 3) - 0.5f;
    }
}
