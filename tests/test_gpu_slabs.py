"""GPU (-m gpu): slab decomposition of ONE grid (BASELINE config 5's scheme, scaled down) must give
the single-device / oracle field bit for bit.  With one visible GPU the slabs share the device
(each slab is still its own context + halo exchange), so the exchange logic is exercised on the
round-end single-GPU box too; tools/slab_multi_gpu.py runs the same check across real devices."""
import numpy as np
import pytest

import oracle
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("axis,slabs", [(0, 2), (0, 4), (1, 3), (2, 2), (0, 8)])
def test_slabs_match_oracle_818(axis, slabs):
    dims = (40, 26, 34)
    v = W.heterogeneous_field(dims, seed=21)
    off = W.star("818")
    start = (33, 5, 30)
    ref, _, _ = oracle.solve(v, off, start)
    tt, st = P.solve_slabs(v, off, start, num_slabs=slabs, slab_axis=axis)
    assert_bit_equal(tt, ref, f"axis={axis} slabs={slabs}")
    assert st.relaxations > 0


def test_slabs_thinner_than_the_halo_and_quirk_node_across_the_interface():
    """Slab thickness 5 < R = 7, and start - o_last = (13,8,8)-(7,1,1) lies in another slab than the start."""
    dims = (20, 17, 13)
    v = W.constant_field(dims)
    off = W.star("818")
    start = (13, 8, 8)
    ref, _, _ = oracle.solve(v, off, start)
    tt, _ = P.solve_slabs(v, off, start, num_slabs=4, slab_axis=0)
    assert_bit_equal(tt, ref)


def test_slabs_equal_single_device_on_a_larger_box():
    dims = (96, 64, 40)
    v = W.contrast_field(dims, seed=3)
    off = W.star("5")
    start = (10, 60, 39)
    one, _ = P.solve(v, off, [start])
    tt, _ = P.solve_slabs(v, off, start, num_slabs=3, slab_axis=0)
    assert_bit_equal(tt, one[0])


def test_slabs_straight_from_a_vbox_file(tmp_path):
    """Every slab loads only its planes (+ ghosts) with the subset reader."""
    dims = (30, 22, 19)
    v = W.heterogeneous_field(dims, seed=8)
    P.vbox_store(tmp_path / "m.vbox", v, origin=(1, 1, 1))
    off, start = W.star("818"), (4, 20, 18)
    ref, _, _ = oracle.solve(v, off, start)
    for axis in (0, 1, 2):
        tt, _ = P.solve_slabs_vbox(tmp_path / "m.vbox", off, start, num_slabs=3, slab_axis=axis)
        assert_bit_equal(tt, ref, f"axis {axis}")


def test_slab_argument_errors():
    v = W.random_field((6, 6, 6))
    with pytest.raises(P.SweepError, match="more slabs"):
        P.solve_slabs(v, W.star("3"), (0, 0, 0), num_slabs=7, slab_axis=0)
    with pytest.raises(P.SweepError, match="outside"):
        P.solve_slabs(v, W.star("3"), (6, 0, 0), num_slabs=2, slab_axis=0)
