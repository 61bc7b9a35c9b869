"""GPU (-m gpu): BASELINE.json full-size configs.  The serial oracle needs ~15 min per source
here, so these compare against hashes produced by the reference's own code in the build
container (tests/golden/full_241.json, tools/make_golden.py full) and use size-independent
properties: the device fixed-point verifier and simple-vs-tiled kernel equality."""
import hashlib
import json

import numpy as np
import pytest

import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

from conftest import ROOT, assert_bit_equal, make_field

pytestmark = pytest.mark.gpu
GOLD = ROOT / "tests" / "golden" / "full_241.json"
GOLD_MORE = ROOT / "tests" / "golden" / "full_241_more.json"   # tools/make_golden.py more (round 2)


def _golden():
    if not GOLD.exists():
        pytest.skip("tests/golden/full_241.json not generated")
    return json.loads(GOLD.read_text())


def test_config2_heterogeneous_818_start4_matches_reference_hashes():
    gold = [g for g in _golden() if g["label"].startswith("config2")]
    v = W.heterogeneous_field((241, 241, 51), seed=7)
    assert hashlib.sha256(v.tobytes()).hexdigest() == gold[0]["v_sha256"]
    starts = [g["start"] for g in gold]
    with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
        ctx.set_model(v); ctx.set_star(W.star("818")); ctx.set_sources(starts)
        st = ctx.run()
        for s, g in enumerate(gold):
            tt = ctx.get_tt(s)
            assert ctx.count_violations(s) == 0
            sample = tt.ravel()[:: g["sample_stride"]].view(np.uint32)
            assert [int(x) for x in sample] == g["sample_bits"], f"source {s}: sampled floats differ"
            assert hashlib.sha256(tt.tobytes()).hexdigest() == g["tt_sha256"], f"source {s}"
    assert st.relaxations > 0


def test_config1_constant_3fs_matches_reference_hash():
    gold = [g for g in _golden() if g["label"].startswith("config1")]
    g = gold[0]
    v = W.constant_field((241, 241, 51))
    tt, st = P.solve(v, W.star("3"), [g["start"]])
    assert hashlib.sha256(tt[0].tobytes()).hexdigest() == g["tt_sha256"]


def test_full_size_simple_and_tiled_agree_and_are_fixed_points():
    v = W.heterogeneous_field((241, 241, 51), seed=7)
    starts = W.starts(111)[[0, 55, 110]]
    a, _ = P.solve(v, W.star("818"), starts, kernel=api.KERNEL_TILED)
    b, _ = P.solve(v, W.star("818"), starts, kernel=api.KERNEL_SIMPLE)
    assert_bit_equal(a, b)
    with P.SweepContext() as ctx:
        ctx.set_model(v); ctx.set_star(W.star("818")); ctx.set_sources(starts)
        ctx.run()
        assert all(ctx.count_violations(s) == 0 for s in range(len(starts)))


def _more(prefix):
    if not GOLD_MORE.exists():
        pytest.skip("tests/golden/full_241_more.json not generated")
    return [g for g in json.loads(GOLD_MORE.read_text()) if g["label"].startswith(prefix)]


def _check(tt, g, what):
    sample = tt.ravel()[:: g["sample_stride"]].view(np.uint32)
    assert [int(x) for x in sample] == g["sample_bits"], f"{what}: sampled floats differ from the reference"
    assert hashlib.sha256(tt.tobytes()).hexdigest() == g["tt_sha256"], what


def test_config3_rows_of_start111_match_reference_hashes_through_waves():
    """Nine rows of docs/start-111 (first, last and seven in between), converged by the reference's own code in the
    build container; here they run as two waves of single-launch solves (9 > 8 sources per launch)."""
    gold = _more("config3_hetero_818_row")
    assert len(gold) >= 9
    v = W.heterogeneous_field((241, 241, 51), seed=7)
    assert hashlib.sha256(v.tobytes()).hexdigest() == gold[0]["v_sha256"]
    s111 = W.starts(111)
    for g in gold:
        row = int(g["label"].rsplit("row", 1)[1])
        assert list(map(int, s111[row])) == g["start"]
    with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
        ctx.set_model(v); ctx.set_star(W.star("818")); ctx.set_sources([g["start"] for g in gold])
        st = ctx.run()
        assert st.relax_launches == 2, "9 sources = 2 waves of single-launch solves"
        for s, g in enumerate(gold):
            assert ctx.count_violations(s) == 0
            _check(ctx.get_tt(s), g, g["label"])


@pytest.mark.parametrize("label", ["full_hetero_5fs_start1", "full_const_5fs_start1"])
def test_full_size_5fs_matches_reference_hash(label):
    gold = _more(label)
    if not gold:
        pytest.skip(f"{label} not in tests/golden/full_241_more.json")
    g = gold[0]
    v = make_field(g["kind"], (241, 241, 51), g["seed"])
    assert hashlib.sha256(v.tobytes()).hexdigest() == g["v_sha256"]
    tt, st = P.solve(v, W.star("5"), [g["start"]], kernel=api.KERNEL_TILED)
    _check(tt[0], g, label)


def test_scaled_config4_like_box_matches_reference_hashes():
    """301x301x64 heterogeneous box (seed 11 like config 4), six bottom-face sources, 818-FS: twice the tiles of the
    241 box, so the six sources run as two waves; the first source again as ONE grid over three parts."""
    gold = _more("config4like_hetero_818_src")
    if not gold:
        pytest.skip("config4-like cases not in tests/golden/full_241_more.json")
    dims = tuple(gold[0]["dims"])
    v = W.heterogeneous_field(dims, seed=11)
    assert hashlib.sha256(v.tobytes()).hexdigest() == gold[0]["v_sha256"]
    tt, st = P.solve(v, W.star("818"), [g["start"] for g in gold], kernel=api.KERNEL_TILED)
    for s, g in enumerate(gold):
        _check(tt[s], g, g["label"])
    one, _ = P.solve_slabs(v, W.star("818"), gold[0]["start"], num_slabs=3, slab_axis=0)
    _check(one, gold[0], gold[0]["label"] + " as one grid over 3 parts")


def test_every_recorded_row_of_config3_matches_the_reference_hash():
    """tests/golden/config3_all.json (tools/make_golden.py config3_all): sha256 of the converged field of every further
    row of docs/start-111 that the reference's own code was run on in the build container.  All of them are solved
    in one go -- waves of single-launch solves, exactly what bench.py times -- and compared."""
    path = ROOT / "tests" / "golden" / "config3_all.json"
    if not path.exists():
        pytest.skip("tests/golden/config3_all.json not generated")
    gold = json.loads(path.read_text())
    if not gold:
        pytest.skip("no rows recorded")
    v = W.heterogeneous_field((241, 241, 51), seed=7)
    assert hashlib.sha256(v.tobytes()).hexdigest() == gold[0]["v_sha256"]
    s111 = W.starts(111)
    rows = [int(g["label"].rsplit("row", 1)[1]) for g in gold]
    for r, g in zip(rows, gold):
        assert list(map(int, s111[r])) == g["start"]
    tt, st = P.solve(v, W.star("818"), s111[rows], kernel=api.KERNEL_TILED)
    bad = [g["label"] for s, g in enumerate(gold) if hashlib.sha256(tt[s].tobytes()).hexdigest() != g["tt_sha256"]]
    assert not bad, f"{len(bad)} of {len(gold)} fields differ from the reference: {bad[:5]}"
