"""ctypes front-end of the CPU oracle (oracle/wrach_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (wrach_b200) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ARITH_UNFUSED = 0
ARITH_SPV = 1
MAX_PARTICLES_IN_CELL = 9


class Settings(ctypes.Structure):
    """config_shader.rs:15-29 — 32 bytes."""
    _fields_ = [
        ("view_dimensions", ctypes.c_float * 2),
        ("view_anchor", ctypes.c_float * 2),
        ("grid_dimensions", ctypes.c_uint32 * 2),
        ("cell_size", ctypes.c_uint32),
        ("particles_in_frame_count", ctypes.c_uint32),
    ]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only; no reference sources needed)."""
    src = os.path.join(_HERE, "wrach_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        sp = ctypes.POINTER(Settings)
        L.wo_cell_coord.restype = ctypes.c_int32
        L.wo_cell_coord.argtypes = [ctypes.c_float, ctypes.c_uint16]
        L.wo_active_grid.restype = None
        L.wo_active_grid.argtypes = [f32p, ctypes.c_uint16, i32p, u32p]
        L.wo_max_particles_per_frame.restype = ctypes.c_uint32
        L.wo_max_particles_per_frame.argtypes = [ctypes.c_uint32, ctypes.c_uint16]
        L.wo_create_packed_data.restype = ctypes.c_uint32
        L.wo_create_packed_data.argtypes = [f32p, ctypes.c_uint16, f32p, ctypes.c_uint32, u32p, f32p, f32p]
        L.wo_cell_key.restype = ctypes.c_uint32
        L.wo_cell_key.argtypes = [sp, ctypes.c_float, ctypes.c_float]
        L.wo_pairs.restype = None
        L.wo_pairs.argtypes = [f32p, ctypes.c_uint32, ctypes.c_int]
        L.wo_k1_physics.restype = None
        L.wo_k1_physics.argtypes = [sp, u32p, f32p, f32p, f32p, f32p, ctypes.c_int]
        L.wo_k2_count.restype = None
        L.wo_k2_count.argtypes = [sp, f32p, u32p]
        L.wo_k3_scan.restype = None
        L.wo_k3_scan.argtypes = [sp, u32p]
        L.wo_k4_pack.restype = None
        L.wo_k4_pack.argtypes = [sp, f32p, f32p, u32p, f32p, f32p]
        L.wo_step.restype = None
        L.wo_step.argtypes = [sp, u32p, f32p, f32p, f32p, f32p, ctypes.c_uint32, ctypes.c_int]
        L.wo_step_parallel.restype = ctypes.c_int
        L.wo_step_parallel.argtypes = [sp, u32p, f32p, f32p, f32p, f32p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        L.wo_max_threads.restype = ctypes.c_int
        L.wo_step_neighbours.restype = ctypes.c_int
        L.wo_step_neighbours.argtypes = [sp, u32p, f32p, f32p, f32p, f32p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        L.wo_generate_scene.restype = None
        L.wo_generate_scene.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_int, f32p]
        _lib = L
    return _lib


def cell_coord(position, cell_size):
    return lib().wo_cell_coord(position, cell_size)


def active_grid(viewport, cell_size):
    bl = np.zeros(2, np.int32)
    grid = np.zeros(2, np.uint32)
    lib().wo_active_grid(np.asarray(viewport, np.float32), cell_size, bl, grid)
    return (int(bl[0]), int(bl[1])), (int(grid[0]), int(grid[1]))


def active_cells(viewport, cell_size):
    """spatial_bin.rs:68-89 as a list of (x, y), row-major."""
    (blx, bly), (gx, gy) = active_grid(viewport, cell_size)
    return [(blx + x, bly + y) for y in range(gy) for x in range(gx)], (gx, gy)


def max_particles_per_frame(total_cells, cell_size):
    return lib().wo_max_particles_per_frame(total_cells, cell_size)


def generate_scene(n, width, height, seed=0x5752414348, first_id=0, pile=False):
    out = np.empty((n, 4), np.float32)
    lib().wo_generate_scene(seed, first_id, n, width, height, int(pile), out.reshape(-1))
    return out


def pairs(positions, arith):
    p = np.ascontiguousarray(positions, np.float32).copy()
    lib().wo_pairs(p.reshape(-1), p.shape[0], arith)
    return p


class OracleWorld:
    """The reference's WrachState + compute worker, CPU only (state.rs:65-101, builder.rs:24-92)."""

    def __init__(self, dimensions, cell_size, arith=ARITH_SPV, capacity=None, neighbours=False):
        self.arith = arith
        self.neighbours = neighbours  # the 3x3 extension (not in the reference, wrach_oracle.h); off = parity mode
        self.dimensions = (int(dimensions[0]), int(dimensions[1]))
        self.cell_size = int(cell_size)
        self.viewport = np.array([0.0, 0.0, self.dimensions[0], self.dimensions[1]], np.float32)
        _, (gx, gy) = active_grid(self.viewport, self.cell_size)
        self.grid = (gx, gy)
        self.cells = gx * gy
        self.total_cells = self.cells + 2  # builder.rs:32-33
        self.capacity = int(capacity) if capacity else max_particles_per_frame(self.cells, self.cell_size)
        self.settings = Settings()
        self.settings.view_dimensions[:] = [float(self.dimensions[0]), float(self.dimensions[1])]
        self.settings.view_anchor[:] = [0.0, 0.0]  # builder.rs:61
        self.settings.grid_dimensions[:] = [gx, gy]
        self.settings.cell_size = self.cell_size
        self.settings.particles_in_frame_count = 0
        self.indices = np.zeros(self.total_cells, np.uint32)
        self.positions_in = np.zeros((self.capacity, 2), np.float32)
        self.velocities_in = np.zeros((self.capacity, 2), np.float32)
        self.positions_out = np.zeros((self.capacity, 2), np.float32)
        self.velocities_out = np.zeros((self.capacity, 2), np.float32)
        self._store = np.zeros((0, 4), np.float32)

    @property
    def n(self):
        return int(self.settings.particles_in_frame_count)

    def pack(self, particles):
        """create_packed_data over `particles` (n,4) -> (indices, positions, velocities)."""
        particles = np.ascontiguousarray(particles, np.float32).reshape(-1, 4)
        n = particles.shape[0]
        indices = np.zeros(self.total_cells, np.uint32)
        pos = np.zeros((max(n, 1), 2), np.float32)
        vel = np.zeros((max(n, 1), 2), np.float32)
        packed = lib().wo_create_packed_data(self.viewport, self.cell_size, particles.reshape(-1), n, indices,
                                             pos.reshape(-1), vel.reshape(-1))
        return indices, pos[:packed], vel[:packed]

    def add_particles(self, particles):
        """state.rs:90-101: the store accumulates, the whole frame is re-packed and re-uploaded."""
        particles = np.ascontiguousarray(particles, np.float32).reshape(-1, 4)
        self._store = np.concatenate([self._store, particles])
        indices, pos, vel = self.pack(self._store)
        n = pos.shape[0]
        if n > self.capacity:
            raise ValueError("more particles than buffer capacity")
        self.indices[:] = indices
        self.positions_in[:n] = pos
        self.velocities_in[:n] = vel
        self.settings.particles_in_frame_count = n

    def step(self, steps=1, threads=1):
        args = (ctypes.byref(self.settings), self.indices, self.positions_in.reshape(-1),
                self.velocities_in.reshape(-1), self.positions_out.reshape(-1), self.velocities_out.reshape(-1),
                steps, self.arith)
        if self.neighbours:
            return lib().wo_step_neighbours(*args, threads)
        if threads == 1:
            lib().wo_step(*args)
            return 1
        return lib().wo_step_parallel(*args, threads)

    def key(self, x, y):
        return lib().wo_cell_key(ctypes.byref(self.settings), x, y)
