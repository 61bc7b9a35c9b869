#!/usr/bin/env python3
"""Long randomised parity sweep (the generator of tests/test_gpu_random.py over many more seeds):
random shapes, fields, stars (shipped / truncated / asymmetric), start points, kernels, loops and
scheduling knobs; every converged field must equal the CPU oracle bit for bit.
usage: random_parity.py FIRST_SEED COUNT"""
import os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api
from test_gpu_random import _case

first = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = 0
t0 = time.time()
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    dims, v, off, starts, label = _case(rng)
    knobs = {"SWEEPTT_BUCKET": rng.choice(["-1", "0.5", "2", "8"]), "SWEEPTT_GROUPS": rng.choice(["1", "2", "3"]),
             "SWEEPTT_PERSIST": rng.choice(["0", "1"]), "SWEEPTT_LOOKAHEAD": rng.choice(["0", "0.03", "4"]),
             "SWEEPTT_INNER": rng.choice(["1", "2", "3"]), "SWEEPTT_TRIGGER_FRAC": rng.choice(["0", "0.4", "0.9"])}
    os.environ.update({k: str(x) for k, x in knobs.items()})
    kernel = int(rng.choice([api.KERNEL_AUTO, api.KERNEL_AUTO, api.KERNEL_SIMPLE]))
    loop = int(rng.choice([api.LOOP_GRAPH, api.LOOP_GRAPH, api.LOOP_BATCHED]))
    tt, st = P.solve(v, off, starts, kernel=kernel, loop=loop)
    for s, p in enumerate(starts):
        ref, _, _ = oracle.solve(v, off, p)
        nd = int((ref.view(np.uint32) != tt[s].view(np.uint32)).sum())
        if nd:
            bad += 1
            print(f"MISMATCH seed {seed}: {label} start={p} knobs={knobs} kernel={kernel} loop={loop}: {nd} floats", flush=True)
    if (seed - first) % 20 == 19:
        print(f"{seed - first + 1} cases, {bad} bad, {time.time() - t0:.0f}s", flush=True)
print("cases", count, "bad", bad)
sys.exit(1 if bad else 0)
