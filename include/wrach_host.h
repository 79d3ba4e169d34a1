/* wrach_host.h — C ABI of the host-side mirror of Wrach's Rust host code for the physics step.
 *
 * The reference's host side is Rust (no toolchain in this image), so it is restated in C++
 * (wrach_b200/csrc/wrach_host.{hpp,cpp}) with the reference's own names, and exported here so any
 * language can drive it.  Citations are relative to the reference tree (tombh/wrach).
 *
 *   WrachConfig            runners/bevy/src/config_app.rs:10-36
 *   SpatialBin             runners/bevy/src/spatial_bin.rs:10-149
 *   ParticleStore          runners/bevy/src/particle_store.rs:14-133
 *   WrachState, GPUUpload  runners/bevy/src/state.rs:17-101
 *   maybe_upload_to_gpu    runners/bevy/src/plugin/build.rs:88-126
 *   tick                   runners/bevy/src/plugin/build.rs:135-158
 *   WrachAPI               runners/api/src/lib.rs:17-87
 */
#ifndef WRACH_HOST_H
#define WRACH_HOST_H
#include <stddef.h>
#include <stdint.h>

#include "wrach_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* config_app.rs:10-36 (default 480 x 352, cell 3) */
typedef struct wrach_config {
    uint16_t dimensions[2];
    uint16_t cell_size;
    uint8_t boundaries_as_dimensions; /* declared by the reference, never read by it */
    uint8_t reserved;
} wrach_config;
void wrach_config_default(wrach_config *out);

/* ---- pure host arithmetic ---------------------------------------------------------------- */
/* SpatialBin::get_cell_coord — spatial_bin.rs:48-64 (f32::div_euclid, then `as i32`) */
int32_t wrach_host_cell_coord(float position, uint16_t cell_size);
/* SpatialBin::get_active_cells — spatial_bin.rs:68-89: first cell and inclusive grid dimensions */
void wrach_host_active_grid(const float viewport[4], uint16_t cell_size, int32_t bottom_left[2], uint32_t grid[2]);
/* ParticleStore::max_particles_per_frame — particle_store.rs:116-133 */
uint32_t wrach_host_max_particles_per_frame(uint32_t total_cells, uint16_t cell_size);

/* Seeded scene of examples/youre-a-pixel.rs:42-58 (see wrach_b200/scene.py): n rows (x, y, vx, vy),
 * x in [x0, x0 + width), y in [0, height) (pile: height * u^4), v in [-0.5, 0.5). */
void wrach_host_generate_scene(uint64_t seed, uint64_t first_id, uint64_t n, float x0, float width, float height,
                               int pile, float *out_xyvv);

/* Self-checks of a packed frame as `tick` reads it back (a whole world, or the strip owning the cell
 * columns [col_begin, col_end) of a grid `grid_x` columns wide); measurement / validation helpers, run
 * on all host threads.  The reference's own tests compare packed frames the same way
 * (runners/bevy/src/compute/03_prefix_sum.rs:151-260, 04_pack_particle_data.rs:73-140).
 * check_packed: 0 = ok; -1 indices not a monotone start table, -2 a position outside the world
 * [0,width]x[0,height], -3 |v| > 1, -4 a particle in a strip that does not own its column, -5 a
 * particle outside the slot range of the cell its position keys to.
 * packed_checksum: order-sensitive sum over particles of hash(global cell, rank in cell, position
 * bits, velocity bits) mod 2^64; strips of one world add up to the whole world's value. */
int wrach_host_check_packed(const uint32_t *indices, uint64_t n_indices, const float *positions, const float *velocities,
                            uint32_t col_begin, uint32_t col_end, uint32_t grid_x, float width, float height,
                            uint16_t cell_size);
uint64_t wrach_host_packed_checksum(const uint32_t *indices, uint64_t n_indices, const float *positions,
                                    const float *velocities, uint32_t col_begin, uint32_t col_end, uint32_t grid_x);

/* ---- WrachState (state.rs) ---------------------------------------------------------------- */
typedef struct wrach_state wrach_state;
wrach_state *wrach_state_new(const wrach_config *config);                      /* state.rs:65-80 */
/* Strip workers (new): a state whose packing covers only the cell columns [col_begin, col_end) of
 * the grid, row-major over those columns; particles elsewhere stay in the store.  shader_settings
 * keeps the GLOBAL grid, which is what wrach_cuda_create_strip / write_settings expect. */
wrach_state *wrach_state_new_strip(const wrach_config *config, uint32_t col_begin, uint32_t col_end);

/* Strips, the collective re-bin of one frame: what strip `rank` does once every
 * strip's row is known.  rows[s * stride + d] = particles strip s sends to strip d (s, d < n_ranks),
 * rows[s * stride + n_ranks] = particle slots of strip s.  Fills send_off[d] (first record of the
 * segment for strip d in this strip's send buffer), recv_off[s] (first record of the segment from
 * strip s in its receive buffer: segments in rank order) and *n_recv (its population after the frame).
 * Returns -1 when every strip fits, else the lowest rank that would outgrow its slots -- every strip
 * computes the same answer from the same rows, so all of them stop together. */
int wrach_host_strip_exchange_plan(uint32_t n_ranks, uint32_t rank, const uint32_t *rows, uint32_t stride,
                                   uint32_t *send_off, uint32_t *recv_off, uint32_t *n_recv);
void wrach_state_free(wrach_state *s);
/* WrachState::add_particles — state.rs:90-101.  particles = n x (x, y, vx, vy).  Queues a
 * GPUUpload::PackedData and a GPUUpload::Settings. */
int wrach_state_add_particles(wrach_state *s, const float *particles_xyvv, uint64_t n);
uint32_t wrach_state_pending_uploads(const wrach_state *s);                    /* gpu_uploads.len() */
void wrach_state_shader_settings(const wrach_state *s, wrach_world_settings *out);
void wrach_state_grid(const wrach_state *s, uint32_t grid[2], uint32_t *total_cells, uint32_t *max_particles);
/* WrachState.packed_data — what `tick` last read back (or nothing before the first tick) */
const uint32_t *wrach_state_packed_indices(const wrach_state *s, uint64_t *len);
const float *wrach_state_packed_positions(const wrach_state *s, uint64_t *len_vec2);
const float *wrach_state_packed_velocities(const wrach_state *s, uint64_t *len_vec2);
/* ParticleStore::create_packed_data on the current store, copied out (for tests / tools).
 * indices needs total_cells entries; positions/velocities need 2 floats per stored particle.
 * Returns particles_in_frame_count. */
uint32_t wrach_state_create_packed_data(wrach_state *s, uint32_t *indices, float *positions, float *velocities);

/* Viewport streaming (the reference declares the pieces -- cells_to_read_from_gpu,
 * a commented-out update_from_gpu, particle_store.rs:22-26,76-85 -- and hard-wires the anchor to 0,
 * builder.rs:61).  update_from_gpu writes what the last tick read back into the store, cell by
 * cell; set_viewport then moves the window (x0, y0, x1, y1) and queues the newly packed frame and
 * its settings for maybe_upload_to_gpu.  The window must keep its grid size (the worker's buffers
 * were created for it) and its anchor must lie on a cell boundary (the store keys by absolute cell,
 * the shaders relative to the anchor); else WRACH_ERR_BAD_ARG. */
int wrach_state_update_from_gpu(wrach_state *s);
/* Sets WrachState.packed_data without a worker (tools, host-only tests, a frame computed elsewhere). */
int wrach_state_set_packed_data(wrach_state *s, const uint32_t *indices, uint64_t n_indices, const float *positions,
                                const float *velocities, uint64_t n_particles);
int wrach_state_set_viewport(wrach_state *s, const float viewport[4]);
uint64_t wrach_state_stored_particles(const wrach_state *s); /* everything in the store, in view or not */

/* ---- the two plugin systems, against a CUDA worker ---------------------------------------- */
int wrach_plugin_maybe_upload_to_gpu(wrach_cuda_worker *worker, wrach_state *s);   /* build.rs:88-126 */
/* tick honours `if !compute_worker.ready() { return; }` (build.rs:139): WRACH_OK = frame read into
 * packed_data, WRACH_TICK_SKIPPED = the worker was still busy, nothing read (as the reference skips
 * the frame).  tick_wait blocks until the enqueued frames are done, then reads -- what the headless
 * WrachAPI::tick needs (runners/api/src/lib.rs:49-52 reads right after app.update()).  Both queue the
 * three read_vec copies and synchronise ONCE; packed_data's storage is page-locked on first use. */
#define WRACH_TICK_SKIPPED 1
int wrach_plugin_tick(wrach_cuda_worker *worker, wrach_state *s);                  /* build.rs:135-158 */
int wrach_plugin_tick_wait(wrach_cuda_worker *worker, wrach_state *s);
/* Same, but only the N live particles are read back (N = last entry of `indices`) instead of the
 * full capacity: SURVEY.md §8f #1.  packed positions / velocities then have length N. */
int wrach_plugin_tick_active(wrach_cuda_worker *worker, wrach_state *s);

/* ---- WrachAPI (runners/api/src/lib.rs) ---------------------------------------------------- */
typedef struct wrach_api wrach_api;
int wrach_api_new(const wrach_config *config, int device, int arith, wrach_api **out); /* lib.rs:30-45 */
void wrach_api_free(wrach_api *a);
int wrach_api_tick(wrach_api *a);                                                   /* lib.rs:49-52 */
int wrach_api_add_particles(wrach_api *a, const float *particles_xyvv, uint64_t n); /* lib.rs:78-81 */
const float *wrach_api_positions(const wrach_api *a, uint64_t *len_vec2);           /* lib.rs:21-22 */
const float *wrach_api_velocities(const wrach_api *a, uint64_t *len_vec2);          /* lib.rs:23-24 */
wrach_state *wrach_api_get_simulation_state(wrach_api *a);                          /* lib.rs:85-87 */
wrach_cuda_worker *wrach_api_worker(wrach_api *a);
const char *wrach_api_last_error(const wrach_api *a);

#ifdef __cplusplus
}
#endif
#endif
