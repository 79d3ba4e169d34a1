"""The opt-in 3x3 neighbour mode (wrach_cuda_set_neighbour_mode) against the checker's extension of
the same name (oracle/wrach_oracle.c: wo_neighbour_pass).  The reference has no such pass -- its
physics only ever looks inside one cell (cell.rs:52-76) -- so nothing
here is a parity claim against the reference: it pins the CUDA path to our own definition, bit for
bit, and checks that the default stays the reference's own-cell physics."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_same_state, f32, make_pair

pytestmark = pytest.mark.gpu


def neighbour_pair(dims, cell, particles, **kw):
    ow, w = make_pair(dims, cell, particles, **kw)
    ow.neighbours = True
    w.set_neighbour_mode(True)
    return ow, w


@pytest.mark.parametrize("arith", [O.ARITH_UNFUSED, O.ARITH_SPV])
@pytest.mark.parametrize("dims,cell,n", [((10, 10), 3, 60), ((64, 48), 3, 2304), ((333, 217), 3, 54000),
                                          ((500, 300), 6, 100000), ((97, 61), 1, 3000)])
def test_uniform_scene_every_frame(arith, dims, cell, n):
    p = O.generate_scene(n, dims[0], dims[1], seed=4321 + n)
    ow, w = neighbour_pair(dims, cell, p, arith=arith, capacity=2 * n + 64)
    for t in range(8):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "frame %d" % (t + 1))


def test_batched_frames_and_openmp_checker():
    n, dims = 200000, (640, 420)
    p = O.generate_scene(n, dims[0], dims[1], seed=78)
    ow, w = neighbour_pair(dims, 3, p)
    ow.step(12, threads=4)
    w.step(5)
    w.step(7)
    assert_same_state(ow, w, "12 frames")
    assert w.stats()["slow_path_steps"] == 0


def test_crowded_cells_only_the_first_nine_take_part():
    """Hundreds per cell in the bottom rows: slots beyond the ninth neither push nor get pushed."""
    n, dims = 60000, (300, 400)
    p = O.generate_scene(n, dims[0], dims[1], seed=12, pile=True)
    ow, w = neighbour_pair(dims, 3, p, capacity=2 * n)
    for t in range(4):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "frame %d" % (t + 1))


def test_far_movers_with_the_mode_on():
    n, dims = 20000, (240, 160)
    p = O.generate_scene(n, dims[0], dims[1], seed=10)
    p[::7, 2:] *= f32(300.0)
    ow, w = neighbour_pair(dims, 3, p)
    ow.step(6)
    w.step(6)
    assert_same_state(ow, w, "6 frames")
    assert w.stats()["slow_path_steps"] >= 1


def test_mode_switches_between_frames_and_defaults_to_the_reference():
    n, dims = 40000, (300, 200)
    p = O.generate_scene(n, dims[0], dims[1], seed=21)
    ow, w = make_pair(dims, 3, p)            # default: the reference's own-cell physics
    ref, w_ref = make_pair(dims, 3, p)
    for on, frames in ((False, 3), (True, 3), (False, 2), (True, 1)):
        ow.neighbours = on
        w.set_neighbour_mode(on)
        ow.step(frames)
        w.step(frames)
        assert_same_state(ow, w, "neighbours=%s" % on)
    ref.step(9)
    w_ref.step(9)
    assert_same_state(ref, w_ref, "default mode")
    assert not np.array_equal(ref.positions_in, ow.positions_in)  # the mode does change the physics


def test_pair_across_a_cell_border_is_pushed_only_with_the_mode_on():
    """Two particles 0.5 apart on either side of the border between cells (0,0) and (1,0)."""
    p = np.array([[2.8, 1.0, 0, 0], [3.3, 1.0, 0, 0]], f32)
    ow, w = make_pair((12, 12), 3, p)
    ow.step(1)
    w.step(1)
    assert_same_state(ow, w)
    assert np.array_equal(ow.positions_in[:2], p[:, :2])   # reference physics: they never meet
    ow2, w2 = neighbour_pair((12, 12), 3, p)
    ow2.step(1)
    w2.step(1)
    assert_same_state(ow2, w2)
    got = ow2.positions_in[:2]
    assert got[0, 0] < p[0, 0] and got[1, 0] > p[1, 0]     # pushed apart, each by its own half
    assert abs(float(got[1, 0] - got[0, 0]) - 1.0) < 1e-5   # ... to MIN_DISTANCE (particles.rs:17)


def test_strip_workers_take_the_mode():
    """Strips exchange ghost columns for the mode (tests/test_gpu_strips.py::test_neighbour_mode_on_strips
    holds them to the checker); here: the call is accepted and can be switched off again."""
    import wrach_b200 as W
    config = W.WrachConfig((120, 60), cell_size=3)
    full = W.WrachState(config)
    _, _, cap = full.grid()
    g = full.shader_settings.copy()
    g.particles_in_frame_count = 0
    w = W.PhysicsComputeWorker(g, 0, cap, strip=(0, 2, None))
    w.set_neighbour_mode(True)
    w.set_neighbour_mode(False)
    w.close()
