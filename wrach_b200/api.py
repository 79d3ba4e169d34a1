"""Python face of the host mirror (include/wrach_host.h), named after the reference's Rust types:
WrachConfig (config_app.rs), WrachState (state.rs), the plugin systems (plugin/build.rs) and
WrachAPI (runners/api/src/lib.rs).  All arithmetic happens in the C++ library."""
import ctypes

import numpy as np

from . import _ffi
from ._ffi import WorldSettings

_P = ctypes.c_void_p
_u64p = ctypes.POINTER(ctypes.c_uint64)


class _Config(ctypes.Structure):
    _fields_ = [("dimensions", ctypes.c_uint16 * 2), ("cell_size", ctypes.c_uint16),
                ("boundaries_as_dimensions", ctypes.c_uint8), ("reserved", ctypes.c_uint8)]


HOST_SYMBOLS = {
    "wrach_config_default": (None, [ctypes.POINTER(_Config)]),
    "wrach_host_cell_coord": (ctypes.c_int32, [ctypes.c_float, ctypes.c_uint16]),
    "wrach_host_active_grid": (None, [ctypes.POINTER(ctypes.c_float), ctypes.c_uint16,
                                      ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint32)]),
    "wrach_host_max_particles_per_frame": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_uint16]),
    "wrach_host_generate_scene": (None, [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_float, ctypes.c_int, _P]),
    "wrach_host_packed_checksum": (ctypes.c_uint64, [_P, ctypes.c_uint64, _P, _P, ctypes.c_uint32, ctypes.c_uint32,
                                                     ctypes.c_uint32]),
    "wrach_host_check_packed": (ctypes.c_int, [_P, ctypes.c_uint64, _P, _P, ctypes.c_uint32, ctypes.c_uint32,
                                               ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_uint16]),
    "wrach_state_new": (_P, [ctypes.POINTER(_Config)]),
    "wrach_state_new_strip": (_P, [ctypes.POINTER(_Config), ctypes.c_uint32, ctypes.c_uint32]),
    "wrach_state_free": (None, [_P]),
    "wrach_state_add_particles": (ctypes.c_int, [_P, _P, ctypes.c_uint64]),
    "wrach_state_pending_uploads": (ctypes.c_uint32, [_P]),
    "wrach_state_shader_settings": (None, [_P, ctypes.POINTER(WorldSettings)]),
    "wrach_state_grid": (None, [_P, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                                ctypes.POINTER(ctypes.c_uint32)]),
    "wrach_state_packed_indices": (_P, [_P, _u64p]),
    "wrach_state_packed_positions": (_P, [_P, _u64p]),
    "wrach_state_packed_velocities": (_P, [_P, _u64p]),
    "wrach_state_create_packed_data": (ctypes.c_uint32, [_P, _P, _P, _P]),
    "wrach_state_update_from_gpu": (ctypes.c_int, [_P]),
    "wrach_state_set_packed_data": (ctypes.c_int, [_P, _P, ctypes.c_uint64, _P, _P, ctypes.c_uint64]),
    "wrach_state_set_viewport": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_float)]),
    "wrach_state_stored_particles": (ctypes.c_uint64, [_P]),
    "wrach_plugin_maybe_upload_to_gpu": (ctypes.c_int, [_P, _P]),
    "wrach_plugin_tick": (ctypes.c_int, [_P, _P]),
    "wrach_plugin_tick_wait": (ctypes.c_int, [_P, _P]),
    "wrach_plugin_tick_active": (ctypes.c_int, [_P, _P]),
    "wrach_api_new": (ctypes.c_int, [ctypes.POINTER(_Config), ctypes.c_int, ctypes.c_int, ctypes.POINTER(_P)]),
    "wrach_api_free": (None, [_P]),
    "wrach_api_tick": (ctypes.c_int, [_P]),
    "wrach_api_add_particles": (ctypes.c_int, [_P, _P, ctypes.c_uint64]),
    "wrach_api_positions": (_P, [_P, _u64p]),
    "wrach_api_velocities": (_P, [_P, _u64p]),
    "wrach_api_get_simulation_state": (_P, [_P]),
    "wrach_api_worker": (_P, [_P]),
    "wrach_api_last_error": (ctypes.c_char_p, [_P]),
    "wrach_host_strip_exchange_plan": (ctypes.c_int, [ctypes.c_uint32, ctypes.c_uint32, _P, ctypes.c_uint32, _P, _P, _P]),
}

_bound = False


def _lib():
    global _bound
    L = _ffi.lib()
    if not _bound:
        for name, (res, args) in HOST_SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return L


def _view(ptr, count, dtype, shape_tail=()):
    if not ptr or count == 0:
        return np.zeros((0,) + shape_tail, dtype)
    n = count * int(np.prod(shape_tail)) if shape_tail else count
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype).reshape((count,) + shape_tail)


class WrachConfig:
    """config_app.rs:10-36"""

    def __init__(self, dimensions=(480, 352), boundaries_as_dimensions=False, cell_size=3):
        self.dimensions = (int(dimensions[0]), int(dimensions[1]))
        self.boundaries_as_dimensions = bool(boundaries_as_dimensions)
        self.cell_size = int(cell_size)

    def _c(self):
        c = _Config()
        c.dimensions[:] = self.dimensions
        c.cell_size = self.cell_size
        c.boundaries_as_dimensions = int(self.boundaries_as_dimensions)
        return c


def strip_exchange_plan(rank, rows):
    """wrach_host_strip_exchange_plan: `rows` is the (n_ranks, n_ranks + 1) count matrix of the collective
    re-bin (rows[s, d] = particles strip s sends to strip d, rows[s, n_ranks] = slots of strip s).
    Returns (send_off, recv_off, n_recv, first strip over capacity or -1)."""
    rows = np.ascontiguousarray(rows, np.uint32)
    n = rows.shape[0]
    assert rows.shape == (n, n + 1)
    send_off, recv_off = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    n_recv = ctypes.c_uint32()
    over = _lib().wrach_host_strip_exchange_plan(n, rank, rows.ctypes.data, n + 1, send_off.ctypes.data,
                                                 recv_off.ctypes.data, ctypes.addressof(n_recv))
    return send_off, recv_off, n_recv.value, over


def get_cell_coord(position, cell_size):
    """SpatialBin::get_cell_coord for one axis (spatial_bin.rs:48-64)."""
    return _lib().wrach_host_cell_coord(position, cell_size)


def active_grid(viewport, cell_size):
    """First active cell and inclusive grid dimensions (spatial_bin.rs:68-89 without the list)."""
    vp = (ctypes.c_float * 4)(*viewport)
    bl = (ctypes.c_int32 * 2)()
    grid = (ctypes.c_uint32 * 2)()
    _lib().wrach_host_active_grid(vp, cell_size, bl, grid)
    return (bl[0], bl[1]), (grid[0], grid[1])


def get_active_cells(viewport, cell_size):
    """SpatialBin::get_active_cells (spatial_bin.rs:68-89) -> (cells row-major, (gx, gy))."""
    bl, grid = active_grid(viewport, cell_size)
    cells = [(bl[0] + x, bl[1] + y) for y in range(grid[1]) for x in range(grid[0])]
    return cells, grid


def max_particles_per_frame(total_cells, cell_size):
    """ParticleStore::max_particles_per_frame (particle_store.rs:116-133)."""
    return _lib().wrach_host_max_particles_per_frame(total_cells, cell_size)


class WrachState:
    """state.rs:17-101.  Particles are rows (x, y, vx, vy) of a float32 array."""

    def __init__(self, config=None, _handle=None, _owner=None, columns=None):
        self._lib = _lib()
        self._owner = _owner  # keeps a WrachAPI alive when this is its inner state
        if _handle is not None:
            self._h = _handle
            self._owned = False
        else:
            self.config = config or WrachConfig()
            c = self.config._c()
            if columns is None:
                self._h = self._lib.wrach_state_new(ctypes.byref(c))
            else:  # strip worker: pack only the cell columns [begin, end)
                self._h = self._lib.wrach_state_new_strip(ctypes.byref(c), columns[0], columns[1])
            self._owned = True

    def add_particles(self, particles):
        p = np.ascontiguousarray(particles, np.float32).reshape(-1, 4)
        _ffi.check(self._lib.wrach_state_add_particles(self._h, p.ctypes.data, p.shape[0]))

    @property
    def gpu_uploads_pending(self):
        return self._lib.wrach_state_pending_uploads(self._h)

    @property
    def shader_settings(self):
        s = WorldSettings()
        self._lib.wrach_state_shader_settings(self._h, ctypes.byref(s))
        return s

    def grid(self):
        g = (ctypes.c_uint32 * 2)()
        t, m = ctypes.c_uint32(), ctypes.c_uint32()
        self._lib.wrach_state_grid(self._h, g, ctypes.byref(t), ctypes.byref(m))
        return (g[0], g[1]), t.value, m.value

    def update_from_gpu(self):
        """ParticleStore::update_from_gpu (a stub in the reference, particle_store.rs:76-85): write
        what the last tick read back into the store, cell by cell."""
        _ffi.check(self._lib.wrach_state_update_from_gpu(self._h))

    def set_packed_data(self, indices, positions, velocities):
        """WrachState.packed_data = ... (what `tick` normally fills from the worker)."""
        ind = np.ascontiguousarray(indices, np.uint32)
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 2)
        vel = np.ascontiguousarray(velocities, np.float32).reshape(-1, 2)
        _ffi.check(self._lib.wrach_state_set_packed_data(self._h, ind.ctypes.data, ind.size, pos.ctypes.data,
                                                         vel.ctypes.data, pos.shape[0]))

    def set_viewport(self, viewport):
        """Move the simulated window to (x0, y0, x1, y1) and queue the newly packed frame; same grid
        size, anchor on a cell boundary (include/wrach_host.h)."""
        v = (ctypes.c_float * 4)(*[float(x) for x in viewport])
        _ffi.check(self._lib.wrach_state_set_viewport(self._h, v))

    @property
    def stored_particles(self):
        return int(self._lib.wrach_state_stored_particles(self._h))

    def create_packed_data(self):
        """ParticleStore::create_packed_data -> (indices, positions, velocities)."""
        _, total, _ = self.grid()
        indices = np.zeros(total, np.uint32)
        # first call sizes (the store may also hold off-viewport particles), second call fills
        n = self._lib.wrach_state_create_packed_data(self._h, indices.ctypes.data, None, None)
        pos = np.zeros((max(n, 1), 2), np.float32)
        vel = np.zeros((max(n, 1), 2), np.float32)
        self._lib.wrach_state_create_packed_data(self._h, indices.ctypes.data, pos.ctypes.data, vel.ctypes.data)
        return indices, pos[:n], vel[:n]

    @property
    def packed_data(self):
        """What `tick` last read back: (indices, positions (P,2), velocities (P,2)).  These are VIEWS of the
        C++ state's own vectors (capacity-sized, no copy): valid until the next tick / tick_active /
        set_packed_data / close on this state, which may reallocate or free them -- copy what must outlive that."""
        n = ctypes.c_uint64()
        ip = self._lib.wrach_state_packed_indices(self._h, ctypes.byref(n))
        ind = _view(ip, n.value, np.uint32)
        pp = self._lib.wrach_state_packed_positions(self._h, ctypes.byref(n))
        pos = _view(pp, n.value, np.float32, (2,))
        vp = self._lib.wrach_state_packed_velocities(self._h, ctypes.byref(n))
        vel = _view(vp, n.value, np.float32, (2,))
        return ind, pos, vel

    def close(self):
        if self._owned and self._h:
            self._lib.wrach_state_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def maybe_upload_to_gpu(worker, state):
    """plugin/build.rs:88-126 against a PhysicsComputeWorker."""
    _ffi.check(_lib().wrach_plugin_maybe_upload_to_gpu(worker._h, state._h), worker._h)


def tick(worker, state, wait=True):
    """plugin/build.rs:135-158: read the three buffers back into state.packed_data.  wait=False is
    the reference's own behaviour (`if !compute_worker.ready() { return; }`): returns False when the
    worker was still busy and the frame was skipped; wait=True blocks until the frame is there."""
    fn = _lib().wrach_plugin_tick_wait if wait else _lib().wrach_plugin_tick
    return _ffi.check(fn(worker._h, state._h), worker._h) == 0


def tick_active(worker, state):
    """Like tick(), but reads back only the N live particles."""
    _ffi.check(_lib().wrach_plugin_tick_active(worker._h, state._h), worker._h)


class WrachAPI:
    """runners/api/src/lib.rs:17-87."""

    def __init__(self, config=None, device=0, arith=_ffi.ARITH_SPV):
        self._lib = _lib()
        self.config = config or WrachConfig()
        c = self.config._c()
        self._h = _P()
        _ffi.check(self._lib.wrach_api_new(ctypes.byref(c), device, arith, ctypes.byref(self._h)))

    def tick(self):
        _ffi.check(self._lib.wrach_api_tick(self._h), self._lib.wrach_api_worker(self._h))

    def add_particles(self, particles):
        p = np.ascontiguousarray(particles, np.float32).reshape(-1, 4)
        _ffi.check(self._lib.wrach_api_add_particles(self._h, p.ctypes.data, p.shape[0]))

    @property
    def positions(self):
        n = ctypes.c_uint64()
        return _view(self._lib.wrach_api_positions(self._h, ctypes.byref(n)), n.value, np.float32, (2,)).copy()  # (the next tick reuses the storage)

    @property
    def velocities(self):
        n = ctypes.c_uint64()
        return _view(self._lib.wrach_api_velocities(self._h, ctypes.byref(n)), n.value, np.float32, (2,)).copy()

    def get_simulation_state(self):
        return WrachState(_handle=self._lib.wrach_api_get_simulation_state(self._h), _owner=self)

    def close(self):
        if self._h:
            self._lib.wrach_api_free(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
