//! Rust binding of `libwrach_cuda.so`, the B200 (sm_100a) implementation of Wrach's per-frame
//! physics step behind the compute-worker boundary of the Bevy plugin.
//!
//! * [`sys`] -- the C ABI of `include/wrach_cuda.h`, one `extern "C"` item per entry point.
//! * [`CudaPhysicsWorker`] -- the five calls Wrach makes on
//!   `AppComputeWorker<PhysicsComputeWorker>` (`runners/bevy/src/plugin/build.rs:106-120,139-146`),
//!   with the same names and meaning, plus `run` (what bevy_easy_compute's own system does once per
//!   frame, pass order `runners/bevy/src/compute/builder.rs:86-89`).
//! * feature `bevy` -- the worker as a `Resource`, so `maybe_upload_to_gpu` and `tick` keep their
//!   bodies and change only the type of their first parameter (`INTEGRATION.md` section 3).
//!
//! This crate has not been compiled in the image this repository is built in (no Rust toolchain);
//! the other side of the same boundary is exercised there by the C++ mirror of the host code
//! (`wrach_b200/csrc/wrach_host.cpp`) and by `examples/api_smoke.c`.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_void, CStr};

/// The 32-byte uniform: `runners/bevy/src/config_shader.rs:15-29` (`ShaderWorldSettings`),
/// `shaders/shared/src/lib.rs:19-31`, `assets/shaders/types.wgsl:3-15`.  Offsets 0/8/16/24/28.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, bytemuck::Pod, bytemuck::Zeroable)]
pub struct WorldSettings {
    pub view_dimensions: [f32; 2],
    pub view_anchor: [f32; 2],
    pub grid_dimensions: [u32; 2],
    pub cell_size: u32,
    pub particles_in_frame_count: u32,
}
const _: () = assert!(std::mem::size_of::<WorldSettings>() == 32);

/// `enum wrach_buffer`, in the order of `runners/bevy/src/compute/buffers.rs:8-20`.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Buffer {
    WorldSettingsUniform = 0,
    IndicesMain = 1,
    IndicesBlockSums = 2,
    PositionsIn = 3,
    PositionsOut = 4,
    VelocitiesIn = 5,
    VelocitiesOut = 6,
}

impl Buffer {
    /// The buffer names the reference uses as keys (`compute/buffers.rs:8-20`).
    pub fn from_name(name: &str) -> Option<Self> {
        Some(match name {
            "world_config" => Self::WorldSettingsUniform,
            "indices_main" => Self::IndicesMain,
            "indices_block_sums" => Self::IndicesBlockSums,
            "positions_in" => Self::PositionsIn,
            "positions_out" => Self::PositionsOut,
            "velocities_in" => Self::VelocitiesIn,
            "velocities_out" => Self::VelocitiesOut,
            _ => return None,
        })
    }
}

/// `enum wrach_arith`: where the pair push fuses multiply-adds.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Arith {
    /// the Rust source evaluated natively (`shaders/physics` unit-test path)
    Unfused = 0,
    /// the shipped SPIR-V (`assets/shaders/wrach_physics_shaders.spv`): what a GPU run of the reference computes
    Spv = 1,
}

/// `enum wrach_status` (negative values) with the library's message.
#[derive(Debug, Clone, PartialEq, Eq)]
pub struct Error {
    pub status: i32,
    pub message: String,
}
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "wrach_cuda status {}: {}", self.status, self.message)
    }
}
impl std::error::Error for Error {}
pub type Result<T> = std::result::Result<T, Error>;

/// The C ABI, verbatim (`include/wrach_cuda.h`).
pub mod sys {
    use super::*;

    #[repr(C)]
    pub struct wrach_cuda_worker {
        _private: [u8; 0],
    }

    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct wrach_cuda_stats {
        pub steps_completed: u64,
        pub kernel_launches: u64,
        pub slow_path_steps: u64,
        pub halo_bytes_sent: u64,
        pub last_phys_ms: f32,
        pub last_rebin_ms: f32,
        pub phys_launches_last: u32,
        pub rebin_launches_last: u32,
        pub tile_frames: u64,
        pub tile_fallbacks: u64,
        pub tile_packs: u64,
        pub tile_unpacks: u64,
    }

    extern "C" {
        pub fn wrach_cuda_create(settings: *const WorldSettings, total_cells: u32, max_particles: u32, device: c_int,
                                 arith: c_int, out: *mut *mut wrach_cuda_worker) -> c_int;
        pub fn wrach_cuda_create_strip(global_settings: *const WorldSettings, max_particles: u32, device: c_int,
                                       arith: c_int, rank: c_int, n_ranks: c_int, nccl_unique_id: *const c_void,
                                       out: *mut *mut wrach_cuda_worker) -> c_int;
        pub fn wrach_cuda_nccl_unique_id(out_128_bytes: *mut c_void) -> c_int;
        pub fn wrach_cuda_strip_info(w: *const wrach_cuda_worker, col_begin: *mut u32, col_end: *mut u32,
                                     total_cells: *mut u32) -> c_int;
        pub fn wrach_cuda_strip_group_step(workers: *mut *mut wrach_cuda_worker, n: c_int, n_steps: u32) -> c_int;
        pub fn wrach_cuda_strip_columns(grid_x: u32, rank: c_int, n_ranks: c_int, begin: *mut u32, end: *mut u32);
        pub fn wrach_cuda_destroy(w: *mut wrach_cuda_worker);
        pub fn wrach_cuda_write_slice(w: *mut wrach_cuda_worker, buffer: c_int, src: *const c_void, bytes: usize) -> c_int;
        pub fn wrach_cuda_write_settings(w: *mut wrach_cuda_worker, settings: *const WorldSettings) -> c_int;
        pub fn wrach_cuda_step(w: *mut wrach_cuda_worker, n_steps: u32) -> c_int;
        pub fn wrach_cuda_ready(w: *mut wrach_cuda_worker) -> c_int;
        pub fn wrach_cuda_sync(w: *mut wrach_cuda_worker) -> c_int;
        pub fn wrach_cuda_read(w: *mut wrach_cuda_worker, buffer: c_int, dst: *mut c_void, bytes: usize) -> c_int;
        pub fn wrach_cuda_read_async(w: *mut wrach_cuda_worker, buffer: c_int, dst: *mut c_void, bytes: usize) -> c_int;
        pub fn wrach_cuda_buffer_bytes(w: *const wrach_cuda_worker, buffer: c_int) -> usize;
        pub fn wrach_cuda_device_pointer(w: *mut wrach_cuda_worker, buffer: c_int) -> *mut c_void;
        pub fn wrach_cuda_export_buffer_fd(w: *mut wrach_cuda_worker, buffer: c_int, fd: *mut c_int, alloc_bytes: *mut usize) -> c_int;
        pub fn wrach_cuda_settle(w: *mut wrach_cuda_worker) -> c_int;
        pub fn wrach_cuda_selftest_import_fd(device: c_int, fd: c_int, alloc_bytes: usize, dst: *mut c_void, bytes: usize) -> c_int;
        pub fn wrach_cuda_last_error(w: *const wrach_cuda_worker) -> *const c_char;
        pub fn wrach_cuda_alloc_host(bytes: usize) -> *mut c_void;
        pub fn wrach_cuda_free_host(p: *mut c_void);
        pub fn wrach_cuda_host_register(p: *mut c_void, bytes: usize) -> c_int;
        pub fn wrach_cuda_host_unregister(p: *mut c_void) -> c_int;
        pub fn wrach_cuda_set_neighbour_mode(w: *mut wrach_cuda_worker, enabled: c_int) -> c_int;
        pub fn wrach_cuda_step_timed(w: *mut wrach_cuda_worker, n_steps: u32, elapsed_ms: *mut f32) -> c_int;
        pub fn wrach_cuda_step_profiled(w: *mut wrach_cuda_worker, n_steps: u32, phys_ms_total: *mut f32,
                                        rebin_ms_total: *mut f32) -> c_int;
        pub fn wrach_cuda_get_stats(w: *mut wrach_cuda_worker, out: *mut wrach_cuda_stats) -> c_int;
        pub fn wrach_cuda_selftest_push_division(device: c_int, mismatches: *mut u64) -> c_int;
        pub fn wrach_cuda_selftest_push_sqrt(device: c_int, mismatches: *mut u64) -> c_int;
        pub fn wrach_cuda_version() -> *const c_char;
    }
}

fn last_error(w: *const sys::wrach_cuda_worker) -> String {
    // SAFETY: the library returns a NUL-terminated string it owns (valid until the next call on `w`).
    unsafe {
        let p = sys::wrach_cuda_last_error(w);
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

/// `AppComputeWorker<PhysicsComputeWorker>` on a B200: owns the seven device buffers and the frame's
/// kernels.  Errors that are panics in the reference (`expect`, wgpu validation) are `Err` here; the
/// `*_or_panic` twins keep the reference's behaviour for code that is moved over unchanged.
#[cfg_attr(feature = "bevy", derive(bevy_ecs::system::Resource))]
pub struct CudaPhysicsWorker {
    raw: *mut sys::wrach_cuda_worker,
}
// SAFETY: every entry point takes the handle's mutex and binds its device (include/wrach_cuda.h,
// "Thread-safety"): the handle is neither thread-affine nor racy.
unsafe impl Send for CudaPhysicsWorker {}
unsafe impl Sync for CudaPhysicsWorker {}

impl CudaPhysicsWorker {
    /// `PhysicsComputeWorker::build` (`runners/bevy/src/compute/builder.rs:24-92`): `total_cells` is
    /// grid.x * grid.y + 2 (`builder.rs:30-37`), `max_particles` is
    /// `ParticleStore::max_particles_per_frame()`; `settings.particles_in_frame_count` starts at 0
    /// (`builder.rs:63`).
    pub fn new(settings: &WorldSettings, total_cells: u32, max_particles: u32, device: i32, arith: Arith) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        // SAFETY: plain pointers to live values; `raw` receives an owned handle on success.
        let rc = unsafe { sys::wrach_cuda_create(settings, total_cells, max_particles, device, arith as c_int, &mut raw) };
        if rc < 0 {
            return Err(Error { status: rc, message: last_error(std::ptr::null()) });
        }
        Ok(Self { raw })
    }

    fn check(&self, rc: c_int) -> Result<c_int> {
        if rc < 0 { Err(Error { status: rc, message: last_error(self.raw) }) } else { Ok(rc) }
    }

    /// `write_slice(name, &[T])` -- `plugin/build.rs:106,110,114`.  Copies at offset 0, ordered
    /// before the next step; more bytes than the buffer holds is an error (a wgpu validation panic
    /// in the reference).
    pub fn write_slice<T: bytemuck::Pod>(&mut self, buffer: Buffer, data: &[T]) -> Result<()> {
        // SAFETY: `data` is a live slice of plain-old-data; the library copies out of it before returning.
        let rc = unsafe {
            sys::wrach_cuda_write_slice(self.raw, buffer as c_int, data.as_ptr().cast(), std::mem::size_of_val(data))
        };
        self.check(rc).map(|_| ())
    }

    /// `write(Buffers::WORLD_SETTINGS_UNIFORM, &settings)` -- `plugin/build.rs:118-121`.
    pub fn write(&mut self, settings: &WorldSettings) -> Result<()> {
        // SAFETY: pointer to a live 32-byte value.
        let rc = unsafe { sys::wrach_cuda_write_settings(self.raw, settings) };
        self.check(rc).map(|_| ())
    }

    /// One frame: physics -> count -> scan -> pack (`compute/builder.rs:86-89`).  Does not wait.
    pub fn run(&mut self) -> Result<()> {
        self.step(1)
    }
    pub fn step(&mut self, frames: u32) -> Result<()> {
        // SAFETY: handle owned by self.
        let rc = unsafe { sys::wrach_cuda_step(self.raw, frames) };
        self.check(rc).map(|_| ())
    }

    /// `ready()` -- `plugin/build.rs:139`.
    pub fn ready(&self) -> Result<bool> {
        // SAFETY: handle owned by self.
        let rc = unsafe { sys::wrach_cuda_ready(self.raw) };
        self.check(rc).map(|rc| rc == 1)
    }

    /// `read_vec::<T>(name)` -- `plugin/build.rs:144-146`: the whole buffer at capacity
    /// (`runners/api/src/lib.rs:122-124` expects `len == max_particles`).
    pub fn read_vec<T: bytemuck::Pod>(&self, buffer: Buffer) -> Result<Vec<T>> {
        // SAFETY: handle owned by self; `out` is sized to the byte count the library reports.
        unsafe {
            let bytes = sys::wrach_cuda_buffer_bytes(self.raw, buffer as c_int);
            let mut out = vec![T::zeroed(); bytes / std::mem::size_of::<T>()];
            let rc = sys::wrach_cuda_read(self.raw, buffer as c_int, out.as_mut_ptr().cast(), bytes);
            self.check(rc)?;
            Ok(out)
        }
    }

    /// The first `out.len()` elements only -- e.g. the N live particles instead of the capacity
    /// (`wrach_plugin_tick_active` in the C++ mirror).
    pub fn read_into<T: bytemuck::Pod>(&self, buffer: Buffer, out: &mut [T]) -> Result<()> {
        // SAFETY: `out` is a live, exclusively borrowed slice of plain-old-data.
        let rc = unsafe {
            sys::wrach_cuda_read(self.raw, buffer as c_int, out.as_mut_ptr().cast(), std::mem::size_of_val(out))
        };
        self.check(rc).map(|_| ())
    }

    /// `get_buffer(name)` -- `plugin/bind_groups.rs:71,75`: the CUDA device pointer.  Drawing from it
    /// needs CUDA <-> Vulkan external-memory interop.
    pub fn device_pointer(&mut self, buffer: Buffer) -> *mut c_void {
        // SAFETY: handle owned by self.
        unsafe { sys::wrach_cuda_device_pointer(self.raw, buffer as c_int) }
    }

    /// `get_buffer(name)` for the renderer (`plugin/bind_groups.rs:61-83`) as Vulkan external memory:
    /// an opaque POSIX file descriptor of the allocation `PositionsIn` / `VelocitiesIn` lives in, and its
    /// size.  Import it with `VK_KHR_external_memory_fd` (`VkImportMemoryFdInfoKHR`, OPAQUE_FD,
    /// `allocationSize` = the returned size), bind a `VkBuffer` of the buffer's byte size at offset 0, and
    /// wrap it for wgpu (`wgpu::hal` Vulkan: `Device::buffer_from_raw`).  Call `settle()` before drawing.
    pub fn export_buffer_fd(&mut self, buffer: Buffer) -> Result<(std::os::fd::OwnedFd, usize)> {
        use std::os::fd::FromRawFd;
        let (mut fd, mut bytes) = (-1 as c_int, 0usize);
        // SAFETY: pointers to live locals; on success `fd` is a fresh descriptor this process owns.
        let rc = unsafe { sys::wrach_cuda_export_buffer_fd(self.raw, buffer as c_int, &mut fd, &mut bytes) };
        self.check(rc)?;
        // SAFETY: see above.
        Ok((unsafe { std::os::fd::OwnedFd::from_raw_fd(fd) }, bytes))
    }

    /// Frames done and the buffers current in the reference's packed layout, for a reader outside
    /// the library (the renderer drawing from an exported buffer).
    pub fn settle(&mut self) -> Result<()> {
        // SAFETY: handle owned by self.
        let rc = unsafe { sys::wrach_cuda_settle(self.raw) };
        self.check(rc).map(|_| ())
    }

    /// Opt-in 3x3 neighbour pass before every frame -- an extension the reference only announces
    /// (`shaders/physics/src/cell.rs:1-2`); off by default (`include/wrach_cuda.h`).
    pub fn set_neighbour_mode(&mut self, enabled: bool) -> Result<()> {
        // SAFETY: handle owned by self.
        let rc = unsafe { sys::wrach_cuda_set_neighbour_mode(self.raw, enabled as c_int) };
        self.check(rc).map(|_| ())
    }

    pub fn stats(&mut self) -> Result<sys::wrach_cuda_stats> {
        let mut s = sys::wrach_cuda_stats::default();
        // SAFETY: pointer to a live struct of the layout the header declares.
        let rc = unsafe { sys::wrach_cuda_get_stats(self.raw, &mut s) };
        self.check(rc).map(|_| s)
    }

    // -- the reference's panicking surface, for code moved over unchanged ------------------------
    pub fn write_slice_or_panic<T: bytemuck::Pod>(&mut self, name: &str, data: &[T]) {
        let buffer = Buffer::from_name(name).unwrap_or_else(|| panic!("unknown buffer {name}"));
        self.write_slice(buffer, data).unwrap_or_else(|e| panic!("{e}"));
    }
    pub fn read_vec_or_panic<T: bytemuck::Pod>(&self, name: &str) -> Vec<T> {
        let buffer = Buffer::from_name(name).unwrap_or_else(|| panic!("unknown buffer {name}"));
        self.read_vec(buffer).unwrap_or_else(|e| panic!("{e}"))
    }
}

impl Drop for CudaPhysicsWorker {
    fn drop(&mut self) {
        // SAFETY: `raw` came from wrach_cuda_create and is destroyed exactly once.
        unsafe { sys::wrach_cuda_destroy(self.raw) }
    }
}

/// A fixed-length array in `wrach_cuda_alloc_host` memory: page-locked and on the NUMA node next to
/// the GPU, where `read_into` runs at the full PCIe rate (55 GB/s on the B200 box, against 52 for a
/// `Vec` registered afterwards and 17 for a plain `Vec`).  What
/// `WrachState.packed_data` should be made of when `tick` reads every frame.
pub struct PinnedVec<T: bytemuck::Pod> {
    ptr: *mut T,
    len: usize,
}
// SAFETY: owns its allocation; `T: Pod` has no thread affinity.
unsafe impl<T: bytemuck::Pod> Send for PinnedVec<T> {}
unsafe impl<T: bytemuck::Pod> Sync for PinnedVec<T> {}

impl<T: bytemuck::Pod> PinnedVec<T> {
    /// `len` zeroed elements; `None` when the driver cannot page-lock that much (or there is no device).
    pub fn zeroed(len: usize) -> Option<Self> {
        let bytes = len.checked_mul(std::mem::size_of::<T>())?;
        // SAFETY: plain allocation call; the result is checked before use.
        let ptr = unsafe { sys::wrach_cuda_alloc_host(bytes.max(1)) } as *mut T;
        if ptr.is_null() {
            return None;
        }
        // SAFETY: `ptr` points to at least `bytes` writable bytes; all-zero bytes are a valid `T: Pod`.
        unsafe { std::ptr::write_bytes(ptr.cast::<u8>(), 0, bytes) };
        Some(Self { ptr, len })
    }
}
impl<T: bytemuck::Pod> std::ops::Deref for PinnedVec<T> {
    type Target = [T];
    fn deref(&self) -> &[T] {
        // SAFETY: `ptr` is valid for `len` initialised elements for the lifetime of self.
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
}
impl<T: bytemuck::Pod> std::ops::DerefMut for PinnedVec<T> {
    fn deref_mut(&mut self) -> &mut [T] {
        // SAFETY: as above, exclusively borrowed.
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}
impl<T: bytemuck::Pod> Drop for PinnedVec<T> {
    fn drop(&mut self) {
        // SAFETY: allocated by wrach_cuda_alloc_host, freed exactly once.
        unsafe { sys::wrach_cuda_free_host(self.ptr.cast()) }
    }
}

/// Library build tag, e.g. "wrach_cuda sm_100a r2".
pub fn version() -> String {
    // SAFETY: static NUL-terminated string.
    unsafe { CStr::from_ptr(sys::wrach_cuda_version()).to_string_lossy().into_owned() }
}

#[cfg(test)]
mod tests {
    use super::*;

    /// runners/api/src/lib.rs:102-126 at the worker level: needs a B200.
    #[test]
    #[ignore = "needs a CUDA device"]
    fn three_coincident_particles_five_frames() {
        let settings = WorldSettings { view_dimensions: [10.0, 10.0], grid_dimensions: [4, 4], cell_size: 3, ..Default::default() };
        let mut w = CudaPhysicsWorker::new(&settings, 18, 164, 0, Arith::Spv).unwrap();
        let indices: Vec<u32> = { let mut v = vec![0u32; 18]; for i in 2..18 { v[i] = 3; } v };
        w.write_slice(Buffer::IndicesMain, &indices).unwrap();
        w.write_slice(Buffer::PositionsIn, &[[1.0f32, 1.0]; 3]).unwrap();
        w.write_slice(Buffer::VelocitiesIn, &[[0.1f32, 0.1]; 3]).unwrap();
        w.write(&WorldSettings { particles_in_frame_count: 3, ..settings }).unwrap();
        for _ in 0..5 { w.run().unwrap(); }
        let positions: Vec<[f32; 2]> = w.read_vec(Buffer::PositionsIn).unwrap();
        assert_eq!(positions.len(), 164);
        assert_ne!(positions[0], [0.0, 0.0]);
    }
}
