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
