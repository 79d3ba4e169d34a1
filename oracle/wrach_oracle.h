/* wrach_oracle.h — CPU restatement of Wrach's per-frame particle physics step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the timed CPU baseline.  The CUDA library (wrach_b200/csrc) never links,
 * loads or calls it.
 *
 * What it restates (paths relative to the reference tree, tombh/wrach):
 *   K1 physics     shaders/physics/src/lib.rs:41-64, cell.rs:52-131, particles.rs:31-107,
 *                  particle.rs:22-92   (+ FMA placement of assets/shaders/wrach_physics_shaders.spv,
 *                  see tests/golden/spv_arith.json, produced by oracle/tools/spv_dis.py)
 *   K2 count       assets/shaders/particles_per_cell.wgsl:7-30
 *   K3 scan        assets/shaders/prefix_sum.wgsl:17-123 (semantics: in-place exclusive scan of C+2 items)
 *   K4 pack        assets/shaders/pack_new_particle_data.wgsl:10-45
 *   pass order     runners/bevy/src/compute/builder.rs:86-89
 *   host packing   runners/bevy/src/spatial_bin.rs:48-149, particle_store.rs:54-59,116-133
 *   uniform        runners/bevy/src/config_shader.rs:15-29, assets/shaders/types.wgsl:3-15
 *
 * Parity pin: the reference cannot be compiled here (no Rust/wgpu).  The oracle is pinned against
 * every known-answer test the reference holds for this path (tests/golden/reference_kats.json,
 * transcribed with file:line) and against the SPIR-V disassembly fixture.  Beyond those vectors
 * (multi-particle Gauss-Seidel order, overflow, boundaries, multi-step trajectories) parity is
 * UNPINNED by the reference and is defined by this restatement.
 *
 * Third-party arithmetic restated (not under /root/reference): glam 0.25.0 Vec2 ops
 * (component-wise f32; distance = sqrt(dx*dx + dy*dy)), core f32::clamp, std f32::div_euclid.
 */
#ifndef WRACH_ORACLE_H
#define WRACH_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* config_shader.rs:15-29 / types.wgsl:3-15 — 32 bytes, offsets 0/8/16/24/28 */
typedef struct wo_settings {
    float view_dimensions[2];
    float view_anchor[2];
    uint32_t grid_dimensions[2];
    uint32_t cell_size;
    uint32_t particles_in_frame_count;
} wo_settings;

/* Arithmetic variants of K1's pair push. */
enum { WO_ARITH_UNFUSED = 0, /* Rust source evaluated natively (particles.rs unit test path) */
       WO_ARITH_SPV = 1      /* shipped SPIR-V: fma in dist^2 and both position pushes          */ };

#define WO_MAX_PARTICLES_IN_CELL 9 /* cell.rs:21,29-30 with SPATIAL_BIN_CELL_SIZE = 3 */

/* ---- host-side arithmetic (CPU twin used for uploads) ---- */
/* spatial_bin.rs:48-64: position.div_euclid(cell_size as f32) as i32 */
int32_t wo_cell_coord(float position, uint16_t cell_size);
/* spatial_bin.rs:68-89: bottom-left cell, inclusive grid dimensions */
void wo_active_grid(const float viewport[4], uint16_t cell_size, int32_t bottom_left[2], uint32_t grid[2]);
/* particle_store.rs:116-133 */
uint32_t wo_max_particles_per_frame(uint32_t total_cells, uint16_t cell_size);
/* particle_store.rs:54-59 + spatial_bin.rs:103-149: stable counting sort of (x,y,vx,vy) particles
 * into packed order.  indices must hold C+2 entries, positions/velocities up to n float pairs.
 * Returns the number of particles packed (off-viewport cells are skipped, particle_store.rs:214-228). */
uint32_t wo_create_packed_data(const float viewport[4], uint16_t cell_size, const float *particles_xyvv,
                               uint32_t n, uint32_t *indices, float *positions, float *velocities);

/* ---- device-side restatement ---- */
/* particles_per_cell.wgsl:14-27: u32(floor((x - anchor)/f32(cell_size))) row-major key.
 * Out-of-range float->u32 is undefined in the reference; here it saturates and NaN -> 0. */
uint32_t wo_cell_key(const wo_settings *s, float x, float y);

/* particles.rs:62-94 on a private array of `count` (<= 9) positions, in place. */
void wo_pairs(float *pos_xy, uint32_t count, int arith);

void wo_k1_physics(const wo_settings *s, uint32_t *indices, const float *pos_in, const float *vel_in,
                   float *pos_out, float *vel_out, int arith);
void wo_k2_count(const wo_settings *s, const float *pos_out, uint32_t *indices);
void wo_k3_scan(const wo_settings *s, uint32_t *indices);
/* Canonical order: arrival order of the atomicSub = descending particle index,
 * i.e. a stable counting sort (ascending source slot inside each cell). */
void wo_k4_pack(const wo_settings *s, const float *pos_out, const float *vel_out, uint32_t *indices,
                float *pos_in, float *vel_in);

/* K1..K4 `steps` times, single thread.  pos_out/vel_out are scratch of the same capacity. */
void wo_step(const wo_settings *s, uint32_t *indices, float *pos_in, float *vel_in, float *pos_out,
             float *vel_out, uint32_t steps, int arith);

/* Same results as wo_step (bit-identical, checked in tests), OpenMP over cells / particle chunks.
 * threads <= 0 uses omp_get_max_threads().  Returns the thread count used. */
int wo_step_parallel(const wo_settings *s, uint32_t *indices, float *pos_in, float *vel_in, float *pos_out,
                     float *vel_out, uint32_t steps, int arith, int threads);
int wo_max_threads(void);

/* EXTENSION, not in the reference: opt-in 3x3 neighbour search.  Before
 * K1, each of a cell's first nine particles is pushed away (particles.rs:62-94, its own half only)
 * from the first nine particles of the eight surrounding cells, taken at their frame-start
 * positions, cells row-major, slots ascending.  See wrach_oracle.c for the exact order.  Never on
 * for a reference-parity check. */
void wo_neighbour_pass(const wo_settings *s, const uint32_t *indices, float *pos_in, float *pos_tmp, int arith,
                       int threads);
int wo_step_neighbours(const wo_settings *s, uint32_t *indices, float *pos_in, float *vel_in, float *pos_out,
                       float *vel_out, uint32_t steps, int arith, int threads);

/* Counter-based scene generator (examples/youre-a-pixel.rs:42-58 made reproducible):
 * x ~ U[0,W), y ~ U[0,H) (or H*u^4 when pile != 0), vx,vy ~ U[-0.5,0.5).  Fills n (x,y,vx,vy). */
void wo_generate_scene(uint64_t seed, uint64_t first_id, uint32_t n, float width, float height, int pile,
                       float *particles_xyvv);

#ifdef __cplusplus
}
#endif
#endif
