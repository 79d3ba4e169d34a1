"""Python face of the CUDA compute worker, named after what it replaces in the reference:
`AppComputeWorker<PhysicsComputeWorker>` (runners/bevy/src/compute/builder.rs:10-92) with the
methods Wrach's plugin calls on it (runners/bevy/src/plugin/build.rs:106-120,139-146)."""
import ctypes

import numpy as np

from . import _ffi
from ._ffi import WorldSettings

# runners/bevy/src/compute/buffers.rs:8-20
class Buffers:
    WORLD_SETTINGS_UNIFORM = "world_config"
    INDICES_MAIN = "indices_main"
    INDICES_BLOCK_SUMS = "indices_block_sums"
    POSITIONS_IN = "positions_in"
    POSITIONS_OUT = "positions_out"
    VELOCITIES_IN = "velocities_in"
    VELOCITIES_OUT = "velocities_out"


_BUFFER_IDS = {
    Buffers.WORLD_SETTINGS_UNIFORM: _ffi.WORLD_SETTINGS_UNIFORM,
    Buffers.INDICES_MAIN: _ffi.INDICES_MAIN,
    Buffers.INDICES_BLOCK_SUMS: _ffi.INDICES_BLOCK_SUMS,
    Buffers.POSITIONS_IN: _ffi.POSITIONS_IN,
    Buffers.POSITIONS_OUT: _ffi.POSITIONS_OUT,
    Buffers.VELOCITIES_IN: _ffi.VELOCITIES_IN,
    Buffers.VELOCITIES_OUT: _ffi.VELOCITIES_OUT,
}
_INDEX_BUFFERS = (Buffers.INDICES_MAIN, Buffers.INDICES_BLOCK_SUMS)


class PhysicsComputeWorker:
    """One CUDA device running the physics -> count -> scan -> pack frame (builder.rs:86-89)."""

    def __init__(self, settings, total_cells, max_particles, device=0, arith=_ffi.ARITH_SPV, strip=None):
        self._lib = _ffi.lib()
        self._h = ctypes.c_void_p()
        self.settings = settings.copy()
        self.total_cells = int(total_cells)
        self.max_particles = int(max_particles)
        self.strip = strip
        if strip is None:
            _ffi.check(self._lib.wrach_cuda_create(ctypes.byref(self.settings), self.total_cells,
                                                   self.max_particles, device, arith, ctypes.byref(self._h)))
        else:
            # `settings` describes the GLOBAL world; unique_id = None makes an in-process strip
            rank, n_ranks, unique_id = strip
            buf = ctypes.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
            _ffi.check(self._lib.wrach_cuda_create_strip(ctypes.byref(self.settings), self.max_particles, device,
                                                         arith, rank, n_ranks, buf, ctypes.byref(self._h)))
            b, e, t = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
            self._lib.wrach_cuda_strip_info(self._h, ctypes.byref(b), ctypes.byref(e), ctypes.byref(t))
            self.columns = (b.value, e.value)
            self.total_cells = t.value

    # -- AppComputeWorker surface ---------------------------------------------------------------
    def write_slice(self, name, data):
        """build.rs:106-114.  `data`: numpy array (u32 for indices, f32 (n,2) for particle data)."""
        a = np.ascontiguousarray(data)
        _ffi.check(self._lib.wrach_cuda_write_slice(self._h, _BUFFER_IDS[name], a.ctypes.data, a.nbytes), self._h)

    def write(self, name, settings):
        """build.rs:118-121."""
        assert name == Buffers.WORLD_SETTINGS_UNIFORM
        _ffi.check(self._lib.wrach_cuda_write_settings(self._h, ctypes.byref(settings)), self._h)
        self.settings = settings.copy()

    def ready(self):
        """build.rs:139."""
        return _ffi.check(self._lib.wrach_cuda_ready(self._h), self._h) == 1

    def read_vec(self, name, out=None):
        """build.rs:144-146: the whole buffer, capacity-sized (api/src/lib.rs:122-124)."""
        nbytes = self._lib.wrach_cuda_buffer_bytes(self._h, _BUFFER_IDS[name])
        if out is None:
            out = (np.empty(nbytes // 4, np.uint32) if name in _INDEX_BUFFERS
                   else np.empty((nbytes // 8, 2), np.float32))
        assert out.nbytes == nbytes and out.flags["C_CONTIGUOUS"]
        _ffi.check(self._lib.wrach_cuda_read(self._h, _BUFFER_IDS[name], out.ctypes.data, nbytes), self._h)
        return out

    def read_slice(self, name, out):
        """The first out.nbytes bytes of a buffer (wrach_cuda_read with bytes < capacity): what
        tick_active uses to fetch the N live slots instead of the whole capacity."""
        assert out.flags["C_CONTIGUOUS"]
        _ffi.check(self._lib.wrach_cuda_read(self._h, _BUFFER_IDS[name], out.ctypes.data, out.nbytes), self._h)
        return out

    def read_slice_async(self, name, out):
        """read_slice without the final wait (wrach_cuda_read_async): queue the three copies of a
        tick, then sync() once.  `out` must stay alive and untouched until that sync."""
        assert out.flags["C_CONTIGUOUS"]
        _ffi.check(self._lib.wrach_cuda_read_async(self._h, _BUFFER_IDS[name], out.ctypes.data, out.nbytes), self._h)
        return out

    def export_buffer_fd(self, name):
        """get_buffer for a Vulkan / wgpu renderer (bind_groups.rs:61-83): (fd, allocation bytes) of the
        shareable allocation POSITIONS_IN / VELOCITIES_IN lives in (wrach_cuda_export_buffer_fd).  The caller
        owns the descriptor."""
        fd, nbytes = ctypes.c_int(-1), ctypes.c_size_t(0)
        _ffi.check(self._lib.wrach_cuda_export_buffer_fd(self._h, _BUFFER_IDS[name], ctypes.byref(fd), ctypes.byref(nbytes)), self._h)
        return fd.value, nbytes.value

    def settle(self):
        """Frames done and the buffers current in the packed layout, for a reader outside the library."""
        _ffi.check(self._lib.wrach_cuda_settle(self._h), self._h)

    def get_buffer(self, name):
        """bind_groups.rs:71,75 — device pointer (int)."""
        return self._lib.wrach_cuda_device_pointer(self._h, _BUFFER_IDS[name])

    def set_neighbour_mode(self, enabled):
        """Opt-in 3x3 neighbour pass before the physics of every frame -- an extension the reference
        only announces (cell.rs:1-2); see include/wrach_cuda.h.  Off by default."""
        _ffi.check(self._lib.wrach_cuda_set_neighbour_mode(self._h, int(bool(enabled))), self._h)

    # -- the run system --------------------------------------------------------------------------
    def step(self, n_steps=1):
        _ffi.check(self._lib.wrach_cuda_step(self._h, n_steps), self._h)

    def sync(self):
        _ffi.check(self._lib.wrach_cuda_sync(self._h), self._h)

    def step_timed(self, n_steps):
        ms = ctypes.c_float()
        _ffi.check(self._lib.wrach_cuda_step_timed(self._h, n_steps, ctypes.byref(ms)), self._h)
        return ms.value

    def step_profiled(self, n_steps):
        a, b = ctypes.c_float(), ctypes.c_float()
        _ffi.check(self._lib.wrach_cuda_step_profiled(self._h, n_steps, ctypes.byref(a), ctypes.byref(b)), self._h)
        return a.value, b.value

    def stats(self):
        s = _ffi.Stats()
        _ffi.check(self._lib.wrach_cuda_get_stats(self._h, ctypes.byref(s)), self._h)
        return {k: getattr(s, k) for k, _ in s._fields_}

    def close(self):
        if self._h:
            self._lib.wrach_cuda_destroy(self._h)
            self._h = ctypes.c_void_p()

    @staticmethod
    def strip_columns(grid_x, rank, n_ranks):
        """Columns [begin, end) of the global grid owned by `rank` (wrach_cuda_strip_columns)."""
        b, e = ctypes.c_uint32(), ctypes.c_uint32()
        _ffi.lib().wrach_cuda_strip_columns(grid_x, rank, n_ranks, ctypes.byref(b), ctypes.byref(e))
        return b.value, e.value

    @staticmethod
    def nccl_unique_id():
        buf = ctypes.create_string_buffer(128)
        _ffi.check(_ffi.lib().wrach_cuda_nccl_unique_id(buf))
        return buf.raw

    @staticmethod
    def strip_group_step(workers, n_steps=1):
        """In-process strips 0..n-1 of one world, stepped in lockstep (wrach_cuda_strip_group_step)."""
        arr = (ctypes.c_void_p * len(workers))(*[w._h for w in workers])
        _ffi.check(_ffi.lib().wrach_cuda_strip_group_step(arr, len(workers), n_steps), workers[0]._h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
