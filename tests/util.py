"""Shared helpers for the GPU parity tests: build the same world in the oracle and in the CUDA
worker (through the C ABI), step both, compare bit for bit."""
import numpy as np

from oracle import oracle as O

f32 = np.float32


def settings_from_oracle(ow):
    import wrach_b200
    s = wrach_b200.WorldSettings()
    s.view_dimensions[:] = list(ow.settings.view_dimensions)
    s.view_anchor[:] = list(ow.settings.view_anchor)
    s.grid_dimensions[:] = list(ow.settings.grid_dimensions)
    s.cell_size = ow.settings.cell_size
    s.particles_in_frame_count = ow.settings.particles_in_frame_count
    return s


def make_pair(dims, cell, particles, arith=O.ARITH_SPV, capacity=None, device=0):
    """(oracle world, cuda worker) holding the same uploaded frame."""
    import wrach_b200
    from wrach_b200 import Buffers
    particles = np.ascontiguousarray(particles, f32).reshape(-1, 4)
    ow = O.OracleWorld(dims, cell, arith=arith, capacity=capacity)
    s0 = settings_from_oracle(ow)  # builder.rs:56-66: created with particles_in_frame_count = 0
    w = wrach_b200.PhysicsComputeWorker(s0, ow.total_cells, ow.capacity, device=device, arith=arith)
    ow.add_particles(particles)
    n = ow.n
    # maybe_upload_to_gpu (plugin/build.rs:88-126): PackedData then Settings
    w.write_slice(Buffers.INDICES_MAIN, ow.indices)
    if n:
        w.write_slice(Buffers.POSITIONS_IN, ow.positions_in[:n])
        w.write_slice(Buffers.VELOCITIES_IN, ow.velocities_in[:n])
    w.write(Buffers.WORLD_SETTINGS_UNIFORM, settings_from_oracle(ow))
    return ow, w


def read_state(w):
    from wrach_b200 import Buffers
    return (w.read_vec(Buffers.INDICES_MAIN), w.read_vec(Buffers.POSITIONS_IN), w.read_vec(Buffers.VELOCITIES_IN))


def assert_same_state(ow, w, what=""):
    ind, pos, vel = read_state(w)
    n = ow.n
    assert ind.shape == ow.indices.shape and pos.shape == ow.positions_in.shape, what
    if not np.array_equal(ind, ow.indices):
        bad = np.flatnonzero(ind != ow.indices)
        raise AssertionError("%s indices differ at %d entries, first %d: gpu %d oracle %d" % (
            what, bad.size, bad[0], ind[bad[0]], ow.indices[bad[0]]))
    for name, g, o in (("positions", pos[:n], ow.positions_in[:n]), ("velocities", vel[:n], ow.velocities_in[:n])):
        # bit-exact; a NaN equals a NaN whatever its payload (x86 propagates payloads, the GPU
        # canonicalises them - the reference never looks at them)
        diff = (g.view(np.uint32) != o.view(np.uint32)) & ~(np.isnan(g) & np.isnan(o))
        if diff.any():
            bad = np.flatnonzero(diff.any(axis=1))
            raise AssertionError("%s %s differ at %d slots, first slot %d: gpu %r oracle %r" % (
                what, name, bad.size, bad[0], g[bad[0]], o[bad[0]]))


class WindowedOracle:
    """The reference's intended viewport flow (particle_store.rs:22-26,76-85), restated with numpy and
    the oracle: a store of every particle (rows x, y, vx, vy in insertion order), a window packed
    from it, stepped, written back cell by cell, moved.  The checker for update_from_gpu / set_viewport."""

    def __init__(self, particles, viewport, cell=3, arith=O.ARITH_SPV):
        self.store = np.ascontiguousarray(particles, f32).reshape(-1, 4).copy()
        self.cell, self.arith = cell, arith
        self.set_viewport(viewport)

    def set_viewport(self, viewport):
        self.viewport = np.array(viewport, f32)
        (self.bx, self.by), (self.gx, self.gy) = O.active_grid(self.viewport, self.cell)
        total = self.gx * self.gy + 2
        n_all = self.store.shape[0]
        self.indices = np.zeros(total, np.uint32)
        pos, vel = np.zeros((max(n_all, 1), 2), f32), np.zeros((max(n_all, 1), 2), f32)
        self.n = O.lib().wo_create_packed_data(self.viewport, self.cell, self.store.reshape(-1), n_all, self.indices,
                                                pos.reshape(-1), vel.reshape(-1))
        self.pos, self.vel = pos, vel
        self.settings = O.Settings()
        self.settings.view_dimensions[:] = [float(self.viewport[2] - self.viewport[0]), float(self.viewport[3] - self.viewport[1])]
        self.settings.view_anchor[:] = [float(self.viewport[0]), float(self.viewport[1])]
        self.settings.grid_dimensions[:] = [self.gx, self.gy]
        self.settings.cell_size = self.cell
        self.settings.particles_in_frame_count = self.n

    def step(self, frames):
        import ctypes
        sp, sv = np.zeros_like(self.pos), np.zeros_like(self.vel)
        O.lib().wo_step(ctypes.byref(self.settings), self.indices, self.pos.reshape(-1), self.vel.reshape(-1),
                        sp.reshape(-1), sv.reshape(-1), frames, self.arith)

    def in_window(self, particles):
        cx = np.floor(particles[:, 0] / f32(self.cell)).astype(np.int64) - self.bx
        cy = np.floor(particles[:, 1] / f32(self.cell)).astype(np.int64) - self.by
        return (cx >= 0) & (cx < self.gx) & (cy >= 0) & (cy < self.gy)

    def update_from_gpu(self):
        """Replace the buckets of the window's cells by what the window holds now (packed order)."""
        keep = self.store[~self.in_window(self.store)]
        self.store = np.concatenate([keep, np.concatenate([self.pos[:self.n], self.vel[:self.n]], axis=1)])
