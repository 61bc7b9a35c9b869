"""GPU (-m gpu): the two schedulers of the tiled kernel -- the single persistent launch that builds its
work lists on the device (generations, early builds, busy flags) and the CUDA graph of bulk-synchronous
rounds -- must produce the same bits, for key counts on both sides of the shared-memory snapshot limit
and for every early-build distance."""
import numpy as np
import pytest

import oracle
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def _solve(v, off, starts, monkeypatch, **env):
    for k in ("SWEEPTT_PERSIST", "SWEEPTT_LOOKAHEAD", "SWEEPTT_PERSIST_MAX_KEYS", "SWEEPTT_BUCKET", "SWEEPTT_WAVE"):
        monkeypatch.delenv(k, raising=False)
    for k, val in env.items():
        monkeypatch.setenv(k, str(val))
    with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
        ctx.set_model(v); ctx.set_star(off); ctx.set_sources(starts)
        st = ctx.run()
        tt = [ctx.get_tt(s) for s in range(len(starts))]
        viol = [ctx.count_violations(s) for s in range(len(starts))]
    return tt, st, viol


def test_single_launch_is_one_launch_and_equals_the_graph_of_rounds(monkeypatch):
    v = W.heterogeneous_field((97, 83, 61), seed=5)
    off = W.star("818")
    starts = [(48, 41, 60), (0, 0, 0), (96, 82, 30)]
    a, sa, va = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=1)
    b, sb, vb = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=0)
    assert sa.relax_launches == 1 and sb.relax_launches > 1
    assert va == [0, 0, 0] and vb == [0, 0, 0]
    for s in range(len(starts)):
        assert_bit_equal(a[s], b[s], f"source {s}: single launch vs rounds")
    ref, _, _ = oracle.solve(v[:24, :20, :18].copy(), off, (3, 4, 5))
    c, _, vc = _solve(v[:24, :20, :18].copy(), off, [(3, 4, 5)], monkeypatch, SWEEPTT_PERSIST=1)
    assert vc == [0]
    assert_bit_equal(c[0], ref, "single launch vs oracle")


@pytest.mark.parametrize("look", ["0", "0.02", "0.3", "2", "16"])
def test_every_early_build_distance(look, monkeypatch):
    v = W.contrast_field((70, 64, 40), seed=9)
    off = W.star("5")
    starts = [(10, 60, 39), (69, 0, 0)]
    base, _, vb = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=0)
    got, st, vg = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=1, SWEEPTT_LOOKAHEAD=look)
    assert st.relax_launches == 1 and vb == [0, 0] and vg == [0, 0]
    for s in range(2):
        assert_bit_equal(got[s], base[s], f"lookahead {look}, source {s}")


@pytest.mark.parametrize("name,nsrc", [("3", 3), ("818", 9)])
def test_key_snapshot_in_global_memory_when_the_ring_is_too_small(name, nsrc, monkeypatch):
    # 241x241x51 = 6727 tiles per source: 3 sources exceed the 3-FS kernel's ring (16 Ki words),
    # 9 sources the 818-FS kernel's (53 Ki words); by default such problems run as a graph of rounds
    v = W.heterogeneous_field((241, 241, 51), seed=7)
    off = W.star(name)
    starts = W.starts(111)[:nsrc]
    a, sa, va = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=1, SWEEPTT_PERSIST_MAX_KEYS=4000000, SWEEPTT_WAVE=0)
    b, sb, vb = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=0)
    assert sa.relax_launches == 1 and sb.relax_launches > 1
    assert va == [0] * nsrc and vb == [0] * nsrc
    for s in range(nsrc):
        assert_bit_equal(a[s], b[s], f"{name}-FS source {s}")


def test_more_sources_than_one_launch_holds_run_as_waves_of_single_launches(monkeypatch):
    """5 sources, room for 2 per launch -> waves of 2 + 2 + 1 persistent launches, each on its own slice of the boxes,
    keys and work lists; same bits as the graph of rounds over all sources at once and as one-shot sweeptt_solve
    (whose device->host copies of finished waves run behind the next wave)."""
    v = W.heterogeneous_field((97, 83, 61), seed=5)
    off = W.star("818")
    starts = [(48, 41, 60), (0, 0, 0), (96, 82, 30), (10, 70, 5), (90, 3, 58)]
    with P.SweepContext(kernel=api.KERNEL_TILED) as probe:
        probe.set_model(v); probe.set_star(off); probe.set_sources(starts[:1])
        probe.run()
        ntiles = probe.tiles_per_source
    a, sa, va = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=1, SWEEPTT_PERSIST_MAX_KEYS=2 * ntiles + 1)
    b, sb, vb = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=0, SWEEPTT_WAVE=0)
    assert sa.relax_launches == 3 and sb.relax_launches > 3
    assert va == [0] * 5 and vb == [0] * 5
    for s in range(5):
        assert_bit_equal(a[s], b[s], f"source {s}: waves vs rounds")
    monkeypatch.delenv("SWEEPTT_WAVE", raising=False)
    monkeypatch.setenv("SWEEPTT_PERSIST", "1")
    monkeypatch.setenv("SWEEPTT_PERSIST_MAX_KEYS", str(2 * ntiles + 1))
    c, sc = P.solve(v, off, starts, kernel=api.KERNEL_TILED)
    assert sc.relax_launches == 3 and sc.d2h_bytes == 5 * v.size * 4
    for s in range(5):
        assert_bit_equal(c[s], a[s], f"source {s}: sweeptt_solve vs context")
    P.load_library().sweeptt_release_cache()


def test_problems_beyond_the_key_limit_use_the_graph_of_rounds(monkeypatch):
    v = W.random_field((30, 30, 30), seed=1)
    _, st, viol = _solve(v, W.star("818"), [(1, 2, 3)], monkeypatch, SWEEPTT_PERSIST_MAX_KEYS=8)
    assert st.relax_launches > 1 and viol == [0]


def test_single_launch_schedule_is_nondeterministic_but_the_field_is_not(monkeypatch):
    """The order in which CTAs pop, build and wake is timing dependent; the converged bits must not be
    (tools/stress_persistent.py is the long version of this)."""
    v = W.heterogeneous_field((97, 83, 61), seed=5)
    off = W.star("818")
    starts = [(48, 41, 60), (0, 0, 0)]
    first = None
    for rep in range(6):
        tt, st, viol = _solve(v, off, starts, monkeypatch, SWEEPTT_PERSIST=1,
                              SWEEPTT_LOOKAHEAD=["8", "0", "0.05"][rep % 3], SWEEPTT_BUCKET=["2", "0.5"][rep % 2])
        assert st.relax_launches == 1 and viol == [0, 0]
        if first is None:
            first = tt
        for s in range(2):
            assert_bit_equal(tt[s], first[s], f"repeat {rep}, source {s}")
