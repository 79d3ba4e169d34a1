/* wrach_cuda.h — C ABI of the B200 (sm_100a) compute worker for Wrach's per-frame physics step.
 *
 * This is the drop-in boundary: it replaces what Wrach's Bevy plugin gets from
 * bevy_easy_compute's `AppComputeWorker<PhysicsComputeWorker>` (a third-party crate, 0.15.0, not
 * vendored in the reference).  Every entry point cites the reference call site it stands in for;
 * paths are relative to the reference tree (tombh/wrach).  INTEGRATION.md shows the Rust binding.
 *
 * Plain C: opaque handle, pointers and sizes only.  No exceptions cross this boundary, every call
 * returns a wrach_status (0 = ok, negative = error; wrach_cuda_last_error() has the text).
 * Thread-safety: a handle may be used from any thread, one call at a time per handle (each entry
 * takes the handle's mutex and binds its device) — the contract Bevy's `ResMut<AppComputeWorker>`
 * gives the reference (runners/bevy/src/plugin/build.rs:89,136).
 *
 * There is no CPU fallback: without a CUDA device wrach_cuda_create() fails with WRACH_ERR_CUDA.
 */
#ifndef WRACH_CUDA_H
#define WRACH_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The uniform: runners/bevy/src/config_shader.rs:15-29 == shaders/shared/src/lib.rs:19-31 ==
 * assets/shaders/types.wgsl:3-15.  32 bytes, repr(C), offsets 0/8/16/24/28 (static-asserted). */
typedef struct wrach_world_settings {
    float view_dimensions[2];
    float view_anchor[2];
    uint32_t grid_dimensions[2];
    uint32_t cell_size;
    uint32_t particles_in_frame_count;
} wrach_world_settings;

/* Buffer names: runners/bevy/src/compute/buffers.rs:8-20 (same order as declared in
 * runners/bevy/src/compute/builder.rs:70-84). */
typedef enum wrach_buffer {
    WRACH_WORLD_SETTINGS_UNIFORM = 0, /* "world_config"       32 B                      */
    WRACH_INDICES_MAIN = 1,           /* "indices_main"       u32[total_cells]          */
    WRACH_INDICES_BLOCK_SUMS = 2,     /* "indices_block_sums" u32[total_cells] (scratch) */
    WRACH_POSITIONS_IN = 3,           /* "positions_in"       f32x2[max_particles]      */
    WRACH_POSITIONS_OUT = 4,          /* "positions_out"      f32x2[max_particles]      */
    WRACH_VELOCITIES_IN = 5,          /* "velocities_in"      f32x2[max_particles]      */
    WRACH_VELOCITIES_OUT = 6,         /* "velocities_out"     f32x2[max_particles]      */
    WRACH_BUFFER_COUNT = 7
} wrach_buffer;

typedef enum wrach_status {
    WRACH_OK = 0,
    WRACH_ERR_BAD_ARG = -1,  /* null pointer, unknown buffer, inconsistent settings            */
    WRACH_ERR_CAPACITY = -2, /* bytes > buffer capacity (a wgpu validation panic in the reference) */
    WRACH_ERR_CUDA = -3,     /* CUDA runtime error or no device                                */
    WRACH_ERR_NCCL = -4,     /* NCCL error (strip workers)                                     */
    WRACH_ERR_STATE = -5,    /* call not valid in the worker's current state                   */
    WRACH_ERR_FAR_MIGRATION = -6 /* (round 1: a strip met a particle it could not hand over; no longer returned -- such
                                    frames are finished by the collective re-bin over all strips) */
} wrach_status;

/* Arithmetic variant of the pair push. */
typedef enum wrach_arith {
    WRACH_ARITH_UNFUSED = 0, /* the Rust source evaluated natively (shaders/physics unit-test path) */
    WRACH_ARITH_SPV = 1      /* the shipped SPIR-V: fma in dist^2 and the four position updates     */
} wrach_arith;

typedef struct wrach_cuda_worker wrach_cuda_worker;

/* ---- life cycle ------------------------------------------------------------------------- */

/* PhysicsComputeWorker::build — runners/bevy/src/compute/builder.rs:24-92.
 * Allocates the seven buffers (zero-filled, builder.rs:52-55) and the pass pipeline.
 *   total_cells   = grid.x*grid.y + 2                (builder.rs:30-37, 03_prefix_sum.rs:34-44)
 *   max_particles = capacity of the four particle buffers (particle_store.rs:116-133)
 * Unlike the reference there is no 4 194 304-cell limit (03_prefix_sum.rs:29): the scan is
 * single-pass.  settings->particles_in_frame_count is normally 0 here (builder.rs:63). */
int wrach_cuda_create(const wrach_world_settings *settings, uint32_t total_cells, uint32_t max_particles,
                      int device, int arith, wrach_cuda_worker **out);

/* Strip worker: one of `n_ranks` processes, each owning the cell columns
 * [col_begin, col_end) of the GLOBAL grid described by `global_settings`; after every step the
 * edge columns are exchanged with the neighbouring ranks over NCCL.  `nccl_unique_id` is the
 * 128-byte ncclUniqueId made by rank 0 (wrach_cuda_nccl_unique_id) and shared by the caller
 * (e.g. through torch.distributed); pass NULL for in-process strips stepped with
 * wrach_cuda_strip_group_step.  Uploads and read-backs use the strip's LOCAL packing (row-major over
 * its own columns); wrach_cuda_write_settings takes the GLOBAL grid and the LOCAL particle count.
 * Any displacement is legal, as in the reference (velocities are clamped only after integrating,
 * shaders/physics/src/particles.rs:102-104): a frame in which some particle flies further than one
 * cell -- to any strip -- or more particles leave than an exchange message holds is finished by a
 * collective re-bin over all strips (count matrix, all-to-all of the particles, canonical order), and
 * strips on tile frames go back together to the last uploaded / read-back state and replay up to
 * that frame.  A strip whose population outgrows max_particles reports WRACH_ERR_CAPACITY (create
 * strips with head-room: particles migrate).  Every rank must make the same sequence of calls
 * (uploads, steps, syncs, reads): the strips take collective decisions inside them.  At most 64 strips.
 * New capability — the reference is single-device. */
int wrach_cuda_create_strip(const wrach_world_settings *global_settings, uint32_t max_particles, int device,
                            int arith, int rank, int n_ranks, const void *nccl_unique_id,
                            wrach_cuda_worker **out);
int wrach_cuda_nccl_unique_id(void *out_128_bytes);
/* Columns [begin, end) this worker owns and the size of its (local) indices buffer. */
int wrach_cuda_strip_info(const wrach_cuda_worker *w, uint32_t *col_begin, uint32_t *col_end, uint32_t *total_cells);
/* In-process strips: `workers` are strips 0..n-1 of one world created with nccl_unique_id == NULL
 * (on one or on several devices).  Runs n_steps frames in lockstep and waits for them. */
int wrach_cuda_strip_group_step(wrach_cuda_worker **workers, int n, uint32_t n_steps);
/* Columns [begin,end) of the global grid owned by `rank`: cut on multiples of the library's tile
 * width (22 cell columns) when the grid has at least two tile columns per strip, so that the strips
 * run on the fused tile frames with ghost tile columns; an even split of grid.x otherwise. */
void wrach_cuda_strip_columns(uint32_t grid_x, int rank, int n_ranks, uint32_t *begin, uint32_t *end);

/* Drop of the `AppComputeWorker` resource. */
void wrach_cuda_destroy(wrach_cuda_worker *w);

/* ---- uploads (PreUpdate: maybe_upload_to_gpu, runners/bevy/src/plugin/build.rs:88-126) --- */

/* AppComputeWorker::write_slice(name, &[T]) — build.rs:106,110,114.  Copies `bytes` from `src` to
 * the start of the buffer, ordered before the next step.  bytes > capacity -> WRACH_ERR_CAPACITY.
 * `src` may be reused once the call returns unless it is page-locked memory, in which case it must
 * stay untouched until wrach_cuda_ready() reports 1. */
int wrach_cuda_write_slice(wrach_cuda_worker *w, wrach_buffer buffer, const void *src, size_t bytes);

/* AppComputeWorker::write(WORLD_SETTINGS_UNIFORM, &settings) — build.rs:118-121. */
int wrach_cuda_write_settings(wrach_cuda_worker *w, const wrach_world_settings *settings);

/* ---- the step (the worker's run system: passes in declaration order, builder.rs:86-89) --- */

/* Enqueue `n_steps` frames: physics -> count -> exclusive scan -> pack, each pass seeing the
 * previous one's writes.  Returns without waiting for the device. */
int wrach_cuda_step(wrach_cuda_worker *w, uint32_t n_steps);

/* AppComputeWorker::ready() — build.rs:139.  1 = all enqueued work finished, 0 = still running. */
int wrach_cuda_ready(wrach_cuda_worker *w);

/* Block until ready. */
int wrach_cuda_sync(wrach_cuda_worker *w);

/* ---- read-back (Update: tick, runners/bevy/src/plugin/build.rs:135-158) ------------------ */

/* AppComputeWorker::read_vec::<T>(name) — build.rs:144-146.  Waits for enqueued steps, then copies
 * the first `bytes` of the buffer to `dst`.  Full-capacity reads return the whole buffer, as the
 * reference does (runners/api/src/lib.rs:122-124: len == max_particles). */
int wrach_cuda_read(wrach_cuda_worker *w, wrach_buffer buffer, void *dst, size_t bytes);

/* The same copy without the final wait: three read_vec calls of a `tick` (build.rs:144-146) become
 * three queued copies and ONE wrach_cuda_sync().  `dst` must stay valid until that sync; it only
 * overlaps anything when `dst` is page-locked (wrach_cuda_alloc_host / wrach_cuda_host_register) --
 * into pageable memory the CUDA runtime stages the copy and returns when it is done. */
int wrach_cuda_read_async(wrach_cuda_worker *w, wrach_buffer buffer, void *dst, size_t bytes);

/* Capacity of a buffer in bytes (what read_vec would return). */
size_t wrach_cuda_buffer_bytes(const wrach_cuda_worker *w, wrach_buffer buffer);

/* AppComputeWorker::get_buffer(name) — runners/bevy/src/plugin/bind_groups.rs:71,75 (renderer).
 * Returns the CUDA device pointer (the caller may write through it: the worker re-reads the buffers
 * before the next step). */
void *wrach_cuda_device_pointer(wrach_cuda_worker *w, wrach_buffer buffer);

/* get_buffer for a Vulkan / wgpu renderer — the CUDA half of external-memory
 * interop.  The reference's DrawPlugin binds POSITIONS_IN next to the uniform (bind_groups.rs:61-83);
 * here that buffer lives in CUDA memory, so it is handed over as an opaque POSIX file descriptor:
 * the particle buffer (POSITIONS_IN or VELOCITIES_IN) is moved, once, into a shareable allocation
 * (cuMemCreate with CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) and exported
 * (cuMemExportToShareableHandle).  The renderer imports *fd with VK_KHR_external_memory_fd
 * (VkImportMemoryFdInfoKHR, handle type OPAQUE_FD, allocationSize = *alloc_bytes) and binds a
 * VkBuffer of wrach_cuda_buffer_bytes() at offset 0; it owns the descriptor (one per call).  The
 * buffer holds the reference's packed layout after wrach_cuda_settle().  The Vulkan half cannot be
 * exercised in this image; the tests import the descriptor back into CUDA
 * (cuMemImportFromShareableHandle) and compare bytes. */
int wrach_cuda_export_buffer_fd(wrach_cuda_worker *w, wrach_buffer buffer, int *fd, size_t *alloc_bytes);
/* Wait for the enqueued frames and make the buffers current in the reference's packed layout, for a
 * reader outside the library (a renderer drawing from an exported buffer): what read_vec does before
 * it copies, without the copy.  Unlike wrach_cuda_device_pointer it assumes the reader does not write. */
int wrach_cuda_settle(wrach_cuda_worker *w);
/* Test helper for the export: maps `fd` (as returned above) into this process a second time and copies
 * `bytes` from offset 0 to `dst`.  Consumes the descriptor. */
int wrach_cuda_selftest_import_fd(int device, int fd, size_t alloc_bytes, void *dst, size_t bytes);

/* ---- diagnostics / measurement ----------------------------------------------------------- */

const char *wrach_cuda_last_error(const wrach_cuda_worker *w);

/* Page-locked host memory for callers that want asynchronous copies. */
void *wrach_cuda_alloc_host(size_t bytes);
void wrach_cuda_free_host(void *p);
/* Page-lock / release memory the caller already owns (e.g. the Vec behind WrachState.packed_data),
 * so that read-backs into it run at the full PCIe rate.  0 = ok. */
int wrach_cuda_host_register(void *p, size_t bytes);
int wrach_cuda_host_unregister(void *p);

/* EXTENSION -- no counterpart in the reference.  Its physics module announces "the physics for a
 * cell (and its surroundings)" (shaders/physics/src/cell.rs:1-2) but only ever collides the particles
 * of one cell with each other (cell.rs:52-76, particles.rs:62-83).  With this mode on, every frame
 * starts with a 3x3 neighbour pass: each of a cell's first nine particles (cell.rs:21,29-30) is
 * pushed away -- push_close_particles_apart, particles.rs:85-94, its own half only -- from the
 * first nine particles of the eight surrounding cells, taken at their frame-start positions (cells
 * row-major, slots ascending).  Then the frame proceeds exactly as in the reference.  Off by
 * default; every parity check against the reference runs with it off.  On strip workers (set it on
 * every strip) each frame first exchanges one ghost column per side -- the first-nine positions of
 * the neighbouring strips' edge columns.  Takes effect for frames enqueued after the call. */
int wrach_cuda_set_neighbour_mode(wrach_cuda_worker *w, int enabled);

/* Enqueue n_steps and time them with CUDA events on the worker's own stream (inputs resident,
 * no read-back inside).  Blocks until done. */
int wrach_cuda_step_timed(wrach_cuda_worker *w, uint32_t n_steps, float *elapsed_ms);

/* Same frames, with CUDA events around every kernel: summed device time of the physics kernel and
 * of the re-bin kernel over the n_steps frames (for the per-kernel roofline).  Blocks. */
int wrach_cuda_step_profiled(wrach_cuda_worker *w, uint32_t n_steps, float *phys_ms_total, float *rebin_ms_total);

typedef struct wrach_cuda_stats {
    uint64_t steps_completed;    /* frames finished since create                          */
    uint64_t kernel_launches;    /* kernels of this library launched since create         */
    uint64_t slow_path_steps;    /* frames whose re-bin took the generic (far-mover) path  */
    uint64_t halo_bytes_sent;    /* strip workers: bytes handed to ncclSend since create   */
    float last_phys_ms;          /* per-kernel averages of the last wrach_cuda_step_profiled */
    float last_rebin_ms;
    uint32_t phys_launches_last; /* launches inside the last timed batch                   */
    uint32_t rebin_launches_last;
    uint64_t tile_frames;        /* frames completed by the fused tile kernel (one launch per frame) */
    uint64_t tile_fallbacks;     /* batches the tiles could not hold (far mover / density): replayed on k_phys + k_rebin */
    uint64_t tile_packs;         /* conversions tiles -> packed layout (before a read-back)  */
    uint64_t tile_unpacks;       /* conversions packed layout -> tiles (after an upload)     */
} wrach_cuda_stats;
int wrach_cuda_get_stats(wrach_cuda_worker *w, wrach_cuda_stats *out);

/* Self test (GPU): the hand-written correctly rounded division of the pair push against div.rn over
 * every operand pair the push can produce (6.5e8 divisors); *mismatches must come back 0. */
int wrach_cuda_selftest_push_division(int device, unsigned long long *mismatches);

/* The same for the hand-written correctly rounded square root of the pair push, against sqrt.rn
 * over every squared distance it can be given (2^-100 .. 1 + 2^-22: 8.5e8 values). */
int wrach_cuda_selftest_push_sqrt(int device, unsigned long long *mismatches);

/* Library build tag, e.g. "wrach_cuda sm_100a r1". */
const char *wrach_cuda_version(void);

#ifdef __cplusplus
}
#endif
#endif
