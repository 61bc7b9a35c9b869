"""GPU (-m gpu), needs >= 2 devices (skipped on the single-GPU box): the multi-start dispatcher of sweeptt_solve
over distinct devices (mpi/backup.c:351-363 scheme: sources dealt to devices, no communication) and ONE grid spread
over distinct devices (mpi/16partsmpi.c:740-909 replaced by a shared box in peer memory), bit for bit against the
single-device field and the oracle."""
import numpy as np
import pytest

import oracle
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def _need(n):
    have = P.device_count()
    if have < n:
        pytest.skip(f"needs {n} GPUs, this box has {have}")


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_sources_sharded_over_devices_equal_one_device(ndev):
    _need(ndev)
    v = W.heterogeneous_field((97, 83, 61), seed=5)
    off = W.star("818")
    starts = W.starts(111)[:11] % np.array([97, 83, 61])
    one, s1 = P.solve(v, off, starts, num_devices=1)
    many, sn = P.solve(v, off, starts, num_devices=ndev)
    assert sn.devices_used == ndev and s1.devices_used == 1
    assert_bit_equal(many, one, f"{ndev} devices vs one")
    ref, _, _ = oracle.solve(v[:30, :28, :26].copy(), off, (3, 4, 5))
    got, _ = P.solve(v[:30, :28, :26].copy(), off, [(3, 4, 5), (29, 27, 25), (0, 27, 0)], num_devices=2)
    assert_bit_equal(got[0], ref, "2 devices vs oracle")
    P.load_library().sweeptt_release_cache()


@pytest.mark.parametrize("ndev,axis", [(2, 0), (2, 2), (4, 0), (8, 0)])
def test_one_grid_over_distinct_devices_equals_oracle(ndev, axis):
    _need(ndev)
    dims = (72, 40, 34)
    v = W.heterogeneous_field(dims, seed=21)
    off = W.star("818")
    start = (60, 5, 30)
    one, _ = P.solve(v, off, [start])
    tt, st = P.solve_slabs(v, off, start, num_slabs=ndev, slab_axis=axis)
    assert st.devices_used == ndev
    assert_bit_equal(tt, one[0], f"{ndev} devices, axis {axis}")
    small = v[:24, :20, :18].copy()
    ref, _, _ = oracle.solve(small, off, (20, 4, 5))
    got, _ = P.solve_slabs(small, off, (20, 4, 5), num_slabs=2, slab_axis=0)
    assert_bit_equal(got, ref, "2 devices vs oracle")
