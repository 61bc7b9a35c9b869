"""CPU: the oracle restatement against the reference-generated goldens and (when the build
container's oracle/_ref exists) against the reference's own code sweep by sweep."""
import hashlib

import numpy as np
import pytest

import oracle
from uoparallel_seismic_project_b200 import workloads as W

from conftest import assert_bit_equal, make_field


def test_restatement_matches_reference_goldens(golden_small):
    for m in golden_small:
        v = make_field(m["kind"], m["dims"], m["seed"])
        assert hashlib.sha256(v.tobytes()).digest() == m["v_sha"], "synthetic field generator drifted"
        tt, sweeps, _ = oracle.solve(v, W.star(m["star"]), m["start"])
        assert_bit_equal(tt, m["tt"], f"{m['case']}[{m['idx']}]")
        assert sweeps == m["ref_sweeps"]


@pytest.mark.skipif(oracle.reference() is None, reason="oracle/_ref not built (no /root/reference)")
def test_restatement_sweep_by_sweep_against_reference_code():
    v = W.contrast_field((14, 13, 12), seed=4)
    for name in ("3", "5", "818"):
        off = W.star(name)
        start = (7, 3, 11)
        tt = oracle.init_tt(v.shape, start)
        trace = []
        oracle.ref_solve(v, off, start, per_sweep=lambda k, c, t: trace.append((c, t.copy())))
        for c_ref, t_ref in trace:
            c = oracle.sweep(v, tt, off, start)
            assert c == c_ref
            assert_bit_equal(tt, t_ref, f"{name}-FS per-sweep state")


@pytest.mark.skipif(oracle.reference() is None, reason="oracle/_ref not built")
def test_star_distances_match_reference():
    lib = oracle.reference()
    for name in ("3", "5", "818"):
        off = np.ascontiguousarray(W.star(name), np.int32)
        import ctypes
        assert lib.refh_set_star(ctypes.c_void_p(off.ctypes.data), len(off))
        d_ref = np.array([lib.refh_star_distance(l) for l in range(len(off))], np.float32)
        assert_bit_equal(oracle.star_distances(off), d_ref)


def test_fixed_point_invariant_and_detection(golden_small):
    m = golden_small[3]
    v = make_field(m["kind"], m["dims"], m["seed"])
    off = W.star(m["star"])
    assert oracle.violations(v, m["tt"], off, m["start"]) == 0
    worse = m["tt"].copy()
    worse[2, 3, 4] *= 1.5
    assert oracle.violations(v, worse, off, m["start"]) > 0


def test_work_per_sweep_matches_survey():
    # SURVEY.md §3.1 / BASELINE.md: exact visit counts on the 241x241x51 box
    assert oracle.visits_per_sweep((241, 241, 51), W.star("818")) == 2_246_171_812
    assert oracle.visits_per_sweep((241, 241, 51), W.star("3")) == 278_079_410
    assert W.visits_per_sweep((241, 241, 51), W.star("818")) == 2_246_171_812


def test_quirk_missing_edge_is_real():
    """On a constant field, dropping the start-skip/last-offset quirks changes exactly the node
    start - o_last: the oracle must NOT be 'fixed'."""
    v = W.constant_field((20, 17, 13))
    off = W.star("818")
    start = (12, 9, 11)
    tt, _, _ = oracle.solve(v, off, start)
    n = tuple(np.array(start) - off[-1])
    d = oracle.star_distances(off)
    direct = np.float32(d[-1] * np.float32(0.5)) * np.float32(v[n] + v[start])
    assert tt[n] > direct  # the direct edge {start - o_last, start} is missing from the graph
