"""Oracle self-consistency: OpenMP variant == serial restatement bit for bit, invariants of the step
(count conservation, positions inside the world, |v| <= 1, monotone indices), and the cases the
reference leaves unpinned (overflow > 9, boundary flips, first-step |v| > 1)."""
import numpy as np
import pytest

from oracle import oracle as O

f32 = np.float32


def make_world(dims, cell, n, arith=O.ARITH_SPV, seed=1, vel_scale=1.0, pile=False):
    w = O.OracleWorld(dims, cell, arith=arith, capacity=max(n, 1) * 2 + 64)
    p = O.generate_scene(n, dims[0], dims[1], seed=seed, pile=pile)
    p[:, 2:] *= f32(vel_scale)
    w.add_particles(p)
    return w


def check_invariants(w, n):
    assert w.n == n
    ind = w.indices
    assert ind[0] == 0 and ind[-1] == n
    assert np.all(np.diff(ind.astype(np.int64)) >= 0)
    pos, vel = w.positions_in[:n], w.velocities_in[:n]
    assert np.all(pos[:, 0] >= 0) and np.all(pos[:, 0] <= w.dimensions[0])
    assert np.all(pos[:, 1] >= 0) and np.all(pos[:, 1] <= w.dimensions[1])
    assert np.all(np.abs(vel) <= 1.0)
    # every particle sits in the cell its slot says
    cs = f32(w.cell_size)
    key = (np.floor(pos[:, 1] / cs).astype(np.int64) * w.grid[0] + np.floor(pos[:, 0] / cs).astype(np.int64))
    cell_of_slot = np.searchsorted(ind[1:], np.arange(n), side="right") - 1
    assert np.array_equal(key, cell_of_slot)


@pytest.mark.parametrize("arith", [O.ARITH_UNFUSED, O.ARITH_SPV])
@pytest.mark.parametrize("pile", [False, True])
def test_parallel_equals_serial(arith, pile):
    n = 20000
    a = make_world((200, 150), 3, n, arith=arith, pile=pile)
    b = make_world((200, 150), 3, n, arith=arith, pile=pile)
    for _ in range(6):
        a.step(1, threads=1)
        b.step(1, threads=4)
        assert np.array_equal(a.indices, b.indices)
        assert np.array_equal(a.positions_in, b.positions_in)
        assert np.array_equal(a.velocities_in, b.velocities_in)
    check_invariants(a, n)


def test_invariants_with_wild_first_step_velocities():
    n = 5000
    w = make_world((90, 60), 3, n, vel_scale=400.0)  # |v| up to 200: jumps across the world, clamps after
    for _ in range(4):
        w.step()
        check_invariants(w, n)


def test_overflow_particles_integrate_without_collisions():
    # 14 particles in one cell: slots 0..8 collide, 9..13 only integrate (cell.rs:79-95)
    w = O.OracleWorld((9, 9), 3, capacity=64)
    p = np.zeros((14, 4), f32)
    p[:, 0] = 4.0 + 0.01 * np.arange(14)
    p[:, 1] = 4.5
    p[:, 2] = 0.25
    w.add_particles(p)
    before = w.positions_in[:14].copy()
    O.lib().wo_k1_physics(__import__("ctypes").byref(w.settings), w.indices, w.positions_in.reshape(-1),
                          w.velocities_in.reshape(-1), w.positions_out.reshape(-1),
                          w.velocities_out.reshape(-1), w.arith)
    out = w.positions_out[:14]
    assert np.array_equal(out[9:, 0], before[9:, 0] + f32(0.25))  # untouched by pushes
    assert np.array_equal(out[9:, 1], before[9:, 1])
    assert not np.array_equal(out[:9, 0], before[:9, 0] + f32(0.25))  # pushed apart
    assert np.all(w.indices == 0)  # K1 clears every cell slot and the guard (cell.rs:116-131)


def test_boundary_flip_and_velocity_clamp_order():
    # integrate first, then clamp position with sign flip, then clamp |v| (particles.rs:102-104)
    w = O.OracleWorld((10, 10), 5, capacity=16)
    p = np.array([[9.5, 5.0, 3.0, 0.0],    # overshoots right: x = 10, vx = -3 -> clamped to -1
                  [0.2, 0.1, -0.5, -0.5],  # corner double flip
                  [10.0, 10.0, 0.0, 0.0]], f32)  # exactly on the edge: strict '>' leaves it alone
    w.add_particles(p)
    w.step()
    n = 3
    got = {tuple(np.round(r, 6)) for r in np.concatenate([w.positions_in[:n], w.velocities_in[:n]], 1)}
    assert (10.0, 5.0, -1.0, 0.0) in got
    assert (0.0, 0.0, 0.5, 0.5) in got
    assert (10.0, 10.0, 0.0, 0.0) in got


def test_key_matches_cpu_coord_inside_viewport():
    w = O.OracleWorld((1366, 1024), 3)
    rng = np.random.default_rng(3)
    xs = (rng.random(2000, dtype=f32) * f32(1366)).astype(f32)
    ys = (rng.random(2000, dtype=f32) * f32(1024)).astype(f32)
    for x, y in zip(xs, ys):
        assert w.key(x, y) == O.cell_coord(y, 3) * w.grid[0] + O.cell_coord(x, 3)
    assert w.key(f32("nan"), 0.0) == 0 and w.key(-5.0, 0.0) == 0  # our saturation rule


def test_empty_world_steps():
    w = O.OracleWorld((10, 10), 3)
    w.step(3)
    assert np.all(w.indices == 0)


# ---- the 3x3 neighbour extension (not in the reference; wrach_oracle.h) ------------------------

def test_neighbour_extension_serial_equals_openmp_and_leaves_parity_mode_alone():
    dims, n = (90, 66), 5000
    p = O.generate_scene(n, dims[0], dims[1], seed=31)
    worlds = [O.OracleWorld(dims, 3, neighbours=nb) for nb in (False, True, True)]
    for w in worlds:
        w.add_particles(p)
    worlds[0].step(6)
    worlds[1].step(6)
    worlds[2].step(6, threads=4)
    assert np.array_equal(worlds[1].positions_in, worlds[2].positions_in)
    assert np.array_equal(worlds[1].indices, worlds[2].indices)
    assert not np.array_equal(worlds[0].positions_in, worlds[1].positions_in)
    plain = O.OracleWorld(dims, 3)
    plain.add_particles(p)
    plain.step(6, threads=4)
    assert np.array_equal(plain.positions_in, worlds[0].positions_in)  # the flag defaults to the reference's physics


def test_neighbour_extension_pushes_a_pair_across_a_cell_border():
    p = np.array([[2.8, 1.0, 0, 0], [3.3, 1.0, 0, 0]], np.float32)  # cells (0,0) and (1,0), 0.5 apart
    ref = O.OracleWorld((12, 12), 3)
    ref.add_particles(p)
    ref.step(1)
    assert np.array_equal(ref.positions_in[:2], p[:, :2])  # cell.rs:52-76: own cell only
    nb = O.OracleWorld((12, 12), 3, neighbours=True)
    nb.add_particles(p)
    nb.step(1)
    assert np.allclose(nb.positions_in[:2, 0], [2.55, 3.55], atol=1e-6)  # each moved by its own half: MIN_DISTANCE apart


def test_neighbour_extension_ignores_slots_beyond_the_ninth():
    # twelve particles stacked in cell (0,0); a lone particle just across the border in cell (1,0)
    stack = np.tile(np.array([[2.9, 1.5, 0, 0]], np.float32), (12, 1))
    stack[:, 1] += np.arange(12, dtype=np.float32) * np.float32(0.01)
    lone = np.array([[3.2, 1.5, 0, 0]], np.float32)
    w = O.OracleWorld((12, 12), 3, neighbours=True, capacity=64)
    w.add_particles(np.concatenate([stack, lone]))
    before = w.positions_in[:13].copy()
    import ctypes
    tmp = np.zeros_like(w.positions_in)
    O.lib().wo_neighbour_pass.restype = None
    O.lib().wo_neighbour_pass.argtypes = [ctypes.POINTER(O.Settings), np.ctypeslib.ndpointer(np.uint32),
                                          np.ctypeslib.ndpointer(np.float32), np.ctypeslib.ndpointer(np.float32),
                                          ctypes.c_int, ctypes.c_int]
    O.lib().wo_neighbour_pass(ctypes.byref(w.settings), w.indices, w.positions_in.reshape(-1), tmp.reshape(-1),
                              O.ARITH_SPV, 1)
    after = w.positions_in[:13]
    assert np.array_equal(after[9:12], before[9:12])          # overflow slots (cell.rs:79-95) take no part
    assert (after[:9, 0] < before[:9, 0]).all()               # the first nine were pushed left by the lone one
    assert after[12, 0] > before[12, 0]                        # and it was pushed right by (nine of) them


def test_neighbour_extension_invariants_on_random_small_worlds():
    """Seeded sweep over small and degenerate grids: particle count conserved, indices monotone,
    positions inside the world, |v| <= 1, every particle inside the slot range of the cell it keys
    to, serial == OpenMP -- the same size-independent properties the parity mode is held to."""
    rng = np.random.default_rng(20261017)
    for case in range(24):
        dims = (int(rng.integers(1, 90)), int(rng.integers(1, 70)))
        cell = int(rng.choice([1, 2, 3, 5]))
        n = int(rng.integers(0, 1500))
        p = O.generate_scene(n, dims[0], dims[1], seed=1000 + case) if n else np.zeros((0, 4), np.float32)
        if n:
            p[:, 2:] *= np.float32(rng.choice([0.5, 1.0, 3.0]))
        a = O.OracleWorld(dims, cell, neighbours=True, capacity=n + 64)
        b = O.OracleWorld(dims, cell, neighbours=True, capacity=n + 64)
        if n:
            a.add_particles(p)
            b.add_particles(p)
        a.step(4)
        b.step(4, threads=3)
        assert np.array_equal(a.indices, b.indices) and np.array_equal(a.positions_in, b.positions_in), case
        assert a.n == n and int(a.indices[-1]) == n
        assert (np.diff(a.indices.astype(np.int64)) >= 0).all()
        pos, vel = a.positions_in[:n], a.velocities_in[:n]
        assert (pos[:, 0] >= 0).all() and (pos[:, 0] <= dims[0]).all() and (pos[:, 1] >= 0).all() and (pos[:, 1] <= dims[1]).all()
        assert (np.abs(vel) <= 1).all()
        for i in rng.integers(0, max(n, 1), size=min(n, 50)):
            k = a.key(float(pos[i, 0]), float(pos[i, 1]))
            assert a.indices[k + 1] <= i < a.indices[k + 2]
