"""GPU (-m gpu): randomised parity sweep -- random box shapes, fields, stars (shipped, truncated,
asymmetric), start points, kernels and scheduling knobs -- every field bit-exact against the oracle."""
import numpy as np
import pytest

import oracle
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def _case(rng):
    dims = tuple(int(x) for x in rng.integers(1, [41, 41, 70]))
    kind = rng.choice(["random", "hetero", "contrast", "constant"])
    seed = int(rng.integers(0, 1000))
    v = {"random": lambda: W.random_field(dims, seed), "hetero": lambda: W.heterogeneous_field(dims, seed),
         "contrast": lambda: W.contrast_field(dims, seed), "constant": lambda: W.constant_field(dims, 0.3)}[kind]()
    name = rng.choice(["3", "5", "818"])
    off = W.star(name)
    mode = rng.choice(["full", "prefix", "subset"])
    if mode == "prefix":          # a shorter file: the new last entry becomes the unused one
        off = off[: int(rng.integers(8, len(off)))]
    elif mode == "subset":        # asymmetric star
        keep = np.sort(rng.permutation(len(off))[: int(rng.integers(6, min(len(off), 120)))])
        off = off[keep]
    ns = int(rng.integers(1, 5))
    starts = [tuple(int(rng.integers(0, d)) for d in dims) for _ in range(ns)]
    return dims, v, off, starts, f"{dims} {kind}/{seed} {name}-FS {mode} L={len(off)}"


@pytest.mark.parametrize("seed", range(12))
def test_random_cases(seed, monkeypatch):
    rng = np.random.default_rng(1000 + seed)
    dims, v, off, starts, label = _case(rng)
    knobs = {"SWEEPTT_BUCKET": rng.choice(["-1", "0.5", "2", "8"]), "SWEEPTT_GROUPS": rng.choice(["1", "2", "3"]),
             "SWEEPTT_PERSIST": rng.choice(["0", "1"]), "SWEEPTT_LOOKAHEAD": rng.choice(["0", "0.03", "4"]),
             "SWEEPTT_INNER": rng.choice(["1", "2", "3"])}
    for k, val in knobs.items():
        monkeypatch.setenv(k, str(val))
    kernel = int(rng.choice([api.KERNEL_AUTO, api.KERNEL_SIMPLE]))
    loop = int(rng.choice([api.LOOP_GRAPH, api.LOOP_BATCHED]))
    tt, st = P.solve(v, off, starts, kernel=kernel, loop=loop)
    for s, p in enumerate(starts):
        ref, _, _ = oracle.solve(v, off, p)
        assert_bit_equal(tt[s], ref, f"{label} start={p} knobs={knobs} kernel={kernel} loop={loop}")
