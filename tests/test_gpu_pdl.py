"""Programmatic dependent launch (wrach_worker.cu: launch_frame_kernel) is switched on by the
library only for worlds several waves of blocks long, so the small scenes of the parity suite would
never see it.  Here a cross-section of that suite runs again with the overlap forced on
(WRACH_PDL=2) and forced off (WRACH_PDL=0): same oracle, same bit-exact bar.  The environment is
read when a worker is created, so setting it for the duration of a test is enough."""
import pytest

from oracle import oracle as O
from tests import test_gpu_parity as P
from tests import test_gpu_strips as S

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["2", "0"], ids=["pdl-forced", "pdl-off"])
def pdl(request, monkeypatch):
    monkeypatch.setenv("WRACH_PDL", request.param)
    return request.param


def test_uniform_scenes(pdl):
    P.test_uniform_scene_every_step(O.ARITH_SPV, (333, 217), 3, 54000)
    P.test_uniform_scene_every_step(O.ARITH_UNFUSED, (500, 300), 6, 100000)


def test_batches_and_far_movers(pdl):
    P.test_batched_steps_equal_single_steps(O.ARITH_SPV)
    P.test_wild_first_frame_velocities_take_the_generic_path()
    P.test_far_mover_in_the_middle_of_a_batch()


def test_dense_runs(pdl):
    P.test_pile_skewed_occupancy(O.ARITH_SPV)
    P.test_every_run_over_full()
    P.test_dense_band_next_to_normal_cells()


def test_million_particles_hundred_frames(pdl):
    P.test_config1_one_million_bit_exact()


def test_strips(pdl):
    S.test_strips_equal_single_device_oracle(3, O.ARITH_SPV)
    S.test_strips_many_frames_narrow_world()
    S.test_strips_with_dense_runs(2)
