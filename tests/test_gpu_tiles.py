"""The fused tile frames (wrach_b200/csrc/wrach_tiles.cuh): one launch per frame over 22 x 14-cell
tiles, the state kept tile-major between read-backs.  The whole parity suite already runs through
them (they are the default wherever a scene fits); here are the cases that are about the tiles
themselves -- that they were really used, ragged grids around the tile shape, every way a batch can
fall back to k_phys / k_rebin (and come back), uploads and settings changes between frames -- and a
cross-section of the suite with the tiles switched off (WRACH_TILES=0), which is the path strips,
dense scenes and far movers still take.  Same oracle, same bit-exact bar."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import test_gpu_parity as P
from tests.util import assert_same_state, f32, make_pair, settings_from_oracle

pytestmark = pytest.mark.gpu


def test_uniform_scene_runs_on_tiles_only():
    n, dims = 200000, (640, 420)
    ow, w = make_pair(dims, 3, O.generate_scene(n, dims[0], dims[1], seed=3))
    ow.step(7, threads=4)
    w.step(3)
    w.step(4)
    assert_same_state(ow, w, "7 frames")
    st = w.stats()
    assert st["tile_frames"] == 7 and st["tile_fallbacks"] == 0 and st["slow_path_steps"] == 0
    assert st["tile_unpacks"] == 1 and st["tile_packs"] == 1
    assert st["kernel_launches"] == 1 + 7 + 3  # unpack, seven frames, the three kernels of one pack
    for t in range(3):  # a read-back after every frame: one pack each, never another unpack
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "frame %d" % (8 + t))
    st = w.stats()
    assert st["tile_frames"] == 10 and st["tile_unpacks"] == 1 and st["tile_packs"] == 4


@pytest.mark.parametrize("dims", [(65, 41), (66, 42), (67, 43), (89, 41), (3, 3), (1, 200), (200, 1), (131, 83), (29, 500)])
@pytest.mark.parametrize("arith", [O.ARITH_SPV, O.ARITH_UNFUSED])
def test_grids_around_the_tile_shape(dims, arith):
    """grids of exactly / one less / one more than whole tiles (22 x 14 cells of 3), single rows and columns"""
    n = max(8, int(dims[0] * dims[1] * 0.7))
    p = O.generate_scene(n, dims[0], dims[1], seed=dims[0] * 1000 + dims[1])
    ow, w = make_pair(dims, 3, p, arith=arith, capacity=2 * n + 64)
    for t in range(6):
        ow.step(1)
        w.step(1)
        assert_same_state(ow, w, "step %d" % (t + 1))
    ow.step(9)
    w.step(9)
    assert_same_state(ow, w, "batch of 9")
    st = w.stats()
    assert st["tile_frames"] == 15 and st["tile_fallbacks"] == 0


def test_far_mover_falls_back_and_the_tiles_come_back():
    n, dims = 60000, (420, 300)
    p = O.generate_scene(n, dims[0], dims[1], seed=21)
    p[::11, 2:] *= f32(200.0)  # first frame only: velocities are clamped after it (particles.rs:103-104)
    ow, w = make_pair(dims, 3, p)
    ow.step(30, threads=4)
    w.step(30)  # ONE batch: frame 1 fails on the tiles, is replayed on k_phys / k_rebin (generic re-bin) ...
    assert_same_state(ow, w, "30 frames")
    st = w.stats()
    assert st["steps_completed"] == 30 and st["tile_fallbacks"] == 1 and st["slow_path_steps"] >= 1
    ow.step(12, threads=4)
    w.step(12)  # ... and eight frames later the tiles are tried again
    assert_same_state(ow, w, "42 frames")
    st = w.stats()
    assert st["tile_frames"] >= 12 and st["tile_fallbacks"] == 1


def test_crowding_in_the_middle_of_a_batch():
    """A cell passes 255 particles at frame 3 of a batch: frames 1-2 stay, the batch is packed from
    frame 3's input and replayed from there on the other path, bit for bit."""
    dims = (300, 200)
    bg = O.generate_scene(30000, dims[0], dims[1], seed=33)
    rng = np.random.default_rng(5)
    still = np.zeros((140, 4), f32)
    still[:, 0] = 150.1 + rng.random(140, dtype=f32) * f32(2.8)   # cell column 50: x in [150, 153)
    still[:, 1] = 99.1 + rng.random(140, dtype=f32) * f32(2.8)    # cell row 33
    movers = np.zeros((140, 4), f32)
    movers[:, 0] = f32(155.5)                                     # column 51, three frames from column 50
    movers[:, 1] = 99.1 + rng.random(140, dtype=f32) * f32(2.8)
    movers[:, 2] = f32(-1.0)
    p = np.concatenate([bg, still, movers])
    ow, w = make_pair(dims, 3, p, capacity=2 * len(p))
    ow.step(8)
    w.step(8)
    assert_same_state(ow, w, "8 frames")
    st = w.stats()
    assert st["steps_completed"] == 8 and st["tile_fallbacks"] == 1
    assert 1 <= st["tile_frames"] <= 4, st   # the frames before the crowded one ran on the tiles
    ow.step(3)
    w.step(3)                                # too dense for the tiles: they stay off until the next upload
    assert_same_state(ow, w, "11 frames")
    assert w.stats()["tile_frames"] == st["tile_frames"]


def test_dense_scene_never_enters_the_tiles_and_an_upload_gives_them_another_chance():
    from wrach_b200 import Buffers
    n, dims = 60000, (150, 100)  # 4 per unit area: 36 per cell
    ow, w = make_pair(dims, 3, O.generate_scene(n, dims[0], dims[1], seed=8), capacity=2 * n)
    ow.step(4)
    w.step(4)
    assert_same_state(ow, w, "dense")
    st = w.stats()
    assert st["tile_frames"] == 0 and st["tile_fallbacks"] == 1
    # a sparse frame uploaded into the same worker
    ow2 = O.OracleWorld(dims, 3, capacity=ow.capacity)
    ow2.add_particles(O.generate_scene(9000, dims[0], dims[1], seed=9))
    w.write_slice(Buffers.INDICES_MAIN, ow2.indices)
    w.write_slice(Buffers.POSITIONS_IN, ow2.positions_in[:ow2.n])
    w.write_slice(Buffers.VELOCITIES_IN, ow2.velocities_in[:ow2.n])
    w.write(Buffers.WORLD_SETTINGS_UNIFORM, settings_from_oracle(ow2))
    ow2.step(5)
    w.step(5)
    assert_same_state(ow2, w, "after the upload")
    assert w.stats()["tile_frames"] == 5


def test_partial_upload_between_tile_frames():
    """write_slice of the velocities alone between two batches: applied to the packed layout, the
    tiles are rebuilt from it."""
    from wrach_b200 import Buffers
    n, dims = 50000, (400, 250)
    ow, w = make_pair(dims, 3, O.generate_scene(n, dims[0], dims[1], seed=14))
    ow.step(5)
    w.step(5)
    ow.velocities_in[:ow.n] *= f32(-0.5)
    vel = w.read_vec(Buffers.VELOCITIES_IN)
    vel[:ow.n] *= f32(-0.5)
    w.write_slice(Buffers.VELOCITIES_IN, vel[:ow.n])
    ow.step(5)
    w.step(5)
    assert_same_state(ow, w, "after the partial upload")
    st = w.stats()
    assert st["tile_frames"] == 10 and st["tile_unpacks"] == 2


@pytest.fixture
def tiles_off(monkeypatch):
    monkeypatch.setenv("WRACH_TILES", "0")


def test_cross_section_without_tiles(tiles_off):
    P.test_uniform_scene_every_step(O.ARITH_SPV, (333, 217), 3, 54000)
    P.test_uniform_scene_every_step(O.ARITH_UNFUSED, (500, 300), 6, 100000)
    P.test_batched_steps_equal_single_steps(O.ARITH_SPV)
    P.test_far_mover_in_the_middle_of_a_batch()
    P.test_pile_skewed_occupancy(O.ARITH_SPV)
    P.test_boundaries_corners_and_nan()
    P.test_empty_world_and_ragged_tiles()
    P.test_cell_edges_one_float_either_side((65532, 300), 3)
    P.test_cell_edges_one_float_either_side((4000, 900), 7)


def test_million_particles_hundred_frames_without_tiles(tiles_off):
    P.test_config1_one_million_bit_exact()


def test_sixteen_million_without_tiles(tiles_off):
    P.test_config2_sixteen_million_full_size()
