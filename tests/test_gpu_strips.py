"""Strip workers: the world cut into strips of cell columns, one worker per strip,
edge-column particles exchanged every frame.  Here the strips live in ONE process on ONE device
(wrach_cuda_strip_group_step: same kernels, device-to-device copies instead of NCCL) so the check
runs on a single-GPU box: every cell of every strip must hold exactly what the single-device oracle
holds for that cell, in the same (canonical) order, bit for bit."""
import numpy as np
import pytest

import wrach_b200 as W
from oracle import oracle as O
from tests.util import f32

pytestmark = pytest.mark.gpu


def make_strips(dims, cell, particles, n_strips, arith=O.ARITH_SPV):
    config = W.WrachConfig(dims, cell_size=cell)
    full = W.WrachState(config)
    (gx, gy), _, cap = full.grid()
    gsettings = full.shader_settings.copy()
    gsettings.particles_in_frame_count = 0
    workers, columns = [], []
    for r in range(n_strips):
        cols = W.PhysicsComputeWorker.strip_columns(gx, r, n_strips)
        st = W.WrachState(config, columns=cols)
        st.add_particles(particles)  # the store keeps everything, the packing only this strip's columns
        w = W.PhysicsComputeWorker(gsettings, 0, max(cap, particles.shape[0]), arith=arith, strip=(r, n_strips, None))
        assert w.columns == cols
        W.maybe_upload_to_gpu(w, st)
        workers.append(w)
        columns.append(cols)
    return workers, columns, (gx, gy)


def assert_strips_equal_oracle(workers, columns, grid, ow, what):
    gx, gy = grid
    oind = ow.indices.astype(np.int64)
    total = 0
    for w, (c0, c1) in zip(workers, columns):
        width = c1 - c0
        ind = w.read_vec(W.Buffers.INDICES_MAIN).astype(np.int64)
        pos = w.read_vec(W.Buffers.POSITIONS_IN)
        vel = w.read_vec(W.Buffers.VELOCITIES_IN)
        assert ind.shape[0] == width * gy + 2 and ind[0] == 0
        # per-cell slot ranges, local and global
        lstart = ind[1:-1].reshape(gy, width)
        lend = ind[2:].reshape(gy, width)
        gcell = (np.arange(gy)[:, None] * gx + np.arange(c0, c1)[None, :])
        gstart, gend = oind[1:][gcell], oind[2:][gcell]
        if not np.array_equal(lend - lstart, gend - gstart):
            bad = np.argwhere((lend - lstart) != (gend - gstart))[0]
            raise AssertionError("%s: strip [%d,%d) cell (%d,%d) holds %d particles, oracle %d" % (
                what, c0, c1, c0 + bad[1], bad[0], (lend - lstart)[bad[0], bad[1]], (gend - gstart)[bad[0], bad[1]]))
        n_local = int(ind[-1])
        total += n_local
        # gather the oracle's particles in this strip's local order and compare everything at once
        counts = (gend - gstart).reshape(-1)
        src = np.repeat(gstart.reshape(-1) - np.cumsum(np.concatenate([[0], counts[:-1]])), counts) + np.arange(n_local)
        for name, g, o in (("positions", pos[:n_local], ow.positions_in[src]), ("velocities", vel[:n_local], ow.velocities_in[src])):
            diff = (g.view(np.uint32) != o.view(np.uint32)) & ~(np.isnan(g) & np.isnan(o))
            if diff.any():
                bad = np.flatnonzero(diff.any(axis=1))
                raise AssertionError("%s: strip [%d,%d) %s differ at %d slots, first local slot %d: %r vs %r" % (
                    what, c0, c1, name, bad.size, bad[0], g[bad[0]], o[bad[0]]))
    if len(workers) > 1 or (columns[0][0] == 0 and columns[0][1] == gx):
        assert total == ow.n, "%s: %d particles over all strips, oracle %d" % (what, total, ow.n)


@pytest.mark.parametrize("n_strips", [2, 3, 5])
@pytest.mark.parametrize("arith", [O.ARITH_UNFUSED, O.ARITH_SPV])
def test_strips_equal_single_device_oracle(n_strips, arith):
    dims, n = (400, 260), 78000
    p = O.generate_scene(n, dims[0], dims[1], seed=100 + n_strips)
    ow = O.OracleWorld(dims, 3, arith=arith)
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, n_strips, arith=arith)
    assert_strips_equal_oracle(workers, columns, grid, ow, "upload")
    for t in range(10):
        ow.step(1)
        W.PhysicsComputeWorker.strip_group_step(workers, 1)
        assert_strips_equal_oracle(workers, columns, grid, ow, "frame %d" % (t + 1))
    assert sum(w.stats()["halo_bytes_sent"] for w in workers) > 0


def test_strips_many_frames_narrow_world():
    """Narrow strips (a handful of columns each): most particles cross a boundary sooner or later."""
    dims, n = (60, 300), 13000
    p = O.generate_scene(n, dims[0], dims[1], seed=7)
    ow = O.OracleWorld(dims, 3)
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, 4)
    ow.step(60)
    W.PhysicsComputeWorker.strip_group_step(workers, 25)
    W.PhysicsComputeWorker.strip_group_step(workers, 35)
    assert_strips_equal_oracle(workers, columns, grid, ow, "60 frames")


@pytest.mark.parametrize("n_strips", [2, 3, 5])
def test_far_movers_cross_strips_through_the_collective_rebin(n_strips):
    """A first frame with |v| far above the cell size is legal (the reference clamps velocities only
    after integrating, particles.rs:102-104): particles land several cells -- on narrow strips several
    STRIPS -- away.  Every strip stops its fast re-bin and the collective one (wrach_xrebin.cuh) places
    every particle where the single-device oracle has it, canonical order included."""
    dims, n = (300, 200), 20000
    p = O.generate_scene(n, dims[0], dims[1], seed=3)
    p[:, 2:] *= f32(100.0)  # |v| up to 50: many cells per frame
    ow = O.OracleWorld(dims, 3)
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, n_strips)
    for t in range(3):
        ow.step(1)
        W.PhysicsComputeWorker.strip_group_step(workers, 1)
        assert_strips_equal_oracle(workers, columns, grid, ow, "frame %d" % (t + 1))
    assert all(w.stats()["slow_path_steps"] >= 1 for w in workers)
    ow.step(9)
    W.PhysicsComputeWorker.strip_group_step(workers, 9)
    assert_strips_equal_oracle(workers, columns, grid, ow, "12 frames")
    for w in workers:
        w.close()


def test_more_leavers_than_an_exchange_message_holds():
    """Everybody in the edge columns leaves at once: more than the fixed-size exchange message takes.
    The frame goes to the collective re-bin instead of failing."""
    dims, n = (240, 90), 60000
    rng = np.random.default_rng(8)
    p = np.zeros((n, 4), f32)
    cut = W.PhysicsComputeWorker.strip_columns(dims[0] // 3, 0, 2)[1]
    p[:, 0] = f32(cut * 3 - 1.4) + rng.random(n, dtype=f32) * f32(1.3)   # a band just left of the cut
    p[:, 1] = rng.random(n, dtype=f32) * f32(dims[1])
    p[:, 2] = f32(0.9)
    ow = O.OracleWorld(dims, 3, capacity=2 * n)
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, 2)
    assert columns[0][1] == cut
    for t in range(3):
        ow.step(1)
        W.PhysicsComputeWorker.strip_group_step(workers, 1)
        assert_strips_equal_oracle(workers, columns, grid, ow, "frame %d" % (t + 1))
    assert workers[0].stats()["slow_path_steps"] >= 1
    for w in workers:
        w.close()


@pytest.mark.parametrize("n_strips", [2, 4])
def test_tile_strips_far_mover_in_the_middle_of_a_run(n_strips):
    """Strips on tile frames: a particle that the tiles cannot place (it moves further than one cell)
    sends the frame to k_phys / k_rebin and the collective re-bin on ALL strips, and the tiles come back
    eight frames later."""
    dims, n = (700, 260), 130000   # 234 cell columns: 11 tile columns of 22
    p = O.generate_scene(n, dims[0], dims[1], seed=78)
    p[::13, 2:] *= f32(150.0)      # first frame: wild
    ow = O.OracleWorld(dims, 3)
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, n_strips)
    assert all(c0 % 22 == 0 for c0, _ in columns)
    for upto, k in ((1, 1), (4, 3), (24, 20)):
        ow.step(k)
        W.PhysicsComputeWorker.strip_group_step(workers, k)
        assert_strips_equal_oracle(workers, columns, grid, ow, "%d strips, %d frames" % (n_strips, upto))
    for w in workers:
        st = w.stats()
        assert st["tile_fallbacks"] >= 1 and st["slow_path_steps"] >= 1 and st["tile_frames"] >= 10, st
        w.close()


def test_tile_strips_crowding_in_the_middle_of_a_run():
    """A cell of one strip passes 255 particles after a few tile frames: all the strips leave the tiles
    together and carry on with the particle exchange, bit for bit."""
    dims = (700, 200)
    bg = O.generate_scene(60000, dims[0], dims[1], seed=34)
    rng = np.random.default_rng(6)
    still = np.zeros((140, 4), f32)
    still[:, 0] = 150.1 + rng.random(140, dtype=f32) * f32(2.8)   # cell column 50
    still[:, 1] = 99.1 + rng.random(140, dtype=f32) * f32(2.8)    # cell row 33
    movers = np.zeros((140, 4), f32)
    movers[:, 0] = f32(155.5)                                     # column 51, three frames from column 50
    movers[:, 1] = 99.1 + rng.random(140, dtype=f32) * f32(2.8)
    movers[:, 2] = f32(-1.0)
    p = np.concatenate([bg, still, movers])
    ow = O.OracleWorld(dims, 3, capacity=2 * len(p))
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, 2)
    ow.step(8)
    W.PhysicsComputeWorker.strip_group_step(workers, 8)
    assert_strips_equal_oracle(workers, columns, grid, ow, "8 frames")
    for w in workers:
        st = w.stats()
        assert st["tile_fallbacks"] == 1 and 1 <= st["tile_frames"] <= 4, st
    ow.step(3)
    W.PhysicsComputeWorker.strip_group_step(workers, 3)
    assert_strips_equal_oracle(workers, columns, grid, ow, "11 frames")
    for w in workers:
        w.close()


@pytest.mark.parametrize("n_strips", [2, 3])
@pytest.mark.parametrize("arith", [O.ARITH_UNFUSED, O.ARITH_SPV])
def test_neighbour_mode_on_strips(n_strips, arith):
    """The opt-in 3x3 neighbour mode on strips: the first-nine positions of the neighbours' edge columns
    arrive as ghost columns every frame; result = the single-device checker's, bit for bit."""
    dims, n = (240, 150), 40000
    p = O.generate_scene(n, dims[0], dims[1], seed=55)
    ow = O.OracleWorld(dims, 3, arith=arith, neighbours=True)
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, n_strips, arith=arith)
    for w in workers:
        w.set_neighbour_mode(True)
    for t in range(6):
        ow.step(1)
        W.PhysicsComputeWorker.strip_group_step(workers, 1)
        assert_strips_equal_oracle(workers, columns, grid, ow, "frame %d" % (t + 1))
    ow.step(6)
    W.PhysicsComputeWorker.strip_group_step(workers, 6)
    assert_strips_equal_oracle(workers, columns, grid, ow, "12 frames")
    for w in workers:
        w.close()


@pytest.mark.parametrize("n_strips", [2, 3])
def test_strips_with_dense_runs(n_strips):
    """The pile (hundreds of particles per cell in the bottom rows) cut into strips: dense physics
    mode exports its edge-column leavers, the general re-bin path takes the arrivals as the groups
    of the (absent) neighbour column."""
    dims, n = (360, 300), 90000
    p = O.generate_scene(n, dims[0], dims[1], seed=31, pile=True)
    ow = O.OracleWorld(dims, 3, capacity=2 * n)
    ow.add_particles(p)
    workers, columns, grid = make_strips(dims, 3, p, n_strips)
    for t in range(6):
        ow.step(1)
        W.PhysicsComputeWorker.strip_group_step(workers, 1)
        assert_strips_equal_oracle(workers, columns, grid, ow, "frame %d" % (t + 1))
    ow.step(4)
    W.PhysicsComputeWorker.strip_group_step(workers, 4)
    assert_strips_equal_oracle(workers, columns, grid, ow, "batch of 4 more")


def test_strip_growing_past_its_capacity_is_an_error_not_a_corruption():
    """Arrivals from the neighbouring strip push a strip past the slots it was created with: the frame
    stops before anything is copied and the worker reports WRACH_ERR_CAPACITY (and refuses to go on)."""
    dims, n = (240, 150), 20000
    p = O.generate_scene(n, dims[0], dims[1], seed=17)
    p[:, 2] = np.abs(p[:, 2]) + f32(0.4)  # everybody drifts to the right
    config = W.WrachConfig(dims, cell_size=3)
    full = W.WrachState(config)
    (gx, gy), _, cap = full.grid()
    gsettings = full.shader_settings.copy()
    gsettings.particles_in_frame_count = 0
    workers = []
    for r in range(2):
        cols = W.PhysicsComputeWorker.strip_columns(gx, r, 2)
        st = W.WrachState(config, columns=cols)
        st.add_particles(p)
        n_local = int(st.shader_settings.particles_in_frame_count)
        capacity = n_local + (8 if r == 1 else 4096)  # the right strip has room for eight arrivals only
        w = W.PhysicsComputeWorker(gsettings, 0, capacity, strip=(r, 2, None))
        W.maybe_upload_to_gpu(w, st)
        workers.append(w)
    with pytest.raises(W.WrachCudaError) as e:
        W.PhysicsComputeWorker.strip_group_step(workers, 3)  # (k_phys / k_rebin: found when the frame is re-binned ...
        workers[1].read_vec(W.Buffers.INDICES_MAIN)          #  tile frames: when the tiles are packed for the read-back)
    assert e.value.status == -2 and "grew past" in str(e.value)
    with pytest.raises(W.WrachCudaError) as e2:  # the handle is dead, not silently wrong
        W.PhysicsComputeWorker.strip_group_step(workers, 1)
    assert e2.value.status in (-2, -5)


def test_strips_run_on_tile_frames_when_the_grid_allows():
    """Strips cut on tile boundaries take the fused tile frames (ghost tile columns copied between the
    frames); a grid too narrow for two tile columns per strip keeps k_phys / k_rebin and the particle
    exchange.  Either way: the single-device oracle, bit for bit."""
    dims, n = (700, 260), 130000   # 234 cell columns: 11 tile columns of 22
    p = O.generate_scene(n, dims[0], dims[1], seed=77)
    for n_strips, expect_tiles in ((2, True), (4, True), (7, False)):
        ow = O.OracleWorld(dims, 3)
        ow.add_particles(p)
        workers, columns, grid = make_strips(dims, 3, p, n_strips)
        if expect_tiles:
            assert all(c0 % 22 == 0 for c0, _ in columns)
        for t in range(4):
            ow.step(1)
            W.PhysicsComputeWorker.strip_group_step(workers, 1)
            assert_strips_equal_oracle(workers, columns, grid, ow, "%d strips, frame %d" % (n_strips, t + 1))
        ow.step(16)
        W.PhysicsComputeWorker.strip_group_step(workers, 16)
        assert_strips_equal_oracle(workers, columns, grid, ow, "%d strips, 20 frames" % n_strips)
        for w in workers:
            st = w.stats()
            assert (st["tile_frames"] == 20) == expect_tiles, (n_strips, st)
            w.close()
