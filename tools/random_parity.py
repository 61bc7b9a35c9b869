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
             "SWEEPTT_INNER": rng.choice(["1", "2", "3"]), "SWEEPTT_TRIGGER_FRAC": rng.choice(["0", "0.4", "0.9"]),
             "SWEEPTT_WAVE": rng.choice(["0", "1", "2", "64"]), "SWEEPTT_WAVE_STREAMS": rng.choice(["1", "2"]),
             "SWEEPTT_BLOCK_TILES": rng.choice(["1", "2", "4"])}
    os.environ.update({k: str(x) for k, x in knobs.items()})
    kernel = int(rng.choice([api.KERNEL_AUTO, api.KERNEL_AUTO, api.KERNEL_SIMPLE]))
    loop = int(rng.choice([api.LOOP_GRAPH, api.LOOP_GRAPH, api.LOOP_BATCHED]))
    tt, st = P.solve(v, off, starts, kernel=kernel, loop=loop)
    grid_parts = int(rng.integers(1, 6))
    grid_axis = int(rng.integers(0, 3))
    if seed % 3 == 0 and max(abs(off).max(axis=0)[:2]) <= 7 and abs(off).max() <= 7 and grid_parts <= dims[grid_axis]:
        # ONE grid over several parts (they share the visible devices): same bits as the multi-start path
        try:
            one, _ = P.solve_slabs(v, off, starts[0], num_slabs=grid_parts, slab_axis=grid_axis)
            nd = int((one.view(np.uint32) != tt[0].view(np.uint32)).sum())
        except P.SweepError as e:
            nd = 0 if "tiled kernel" in str(e) else -1
            if nd:
                print("one-grid error:", e, flush=True)
        if nd:
            bad += 1
            print(f"MISMATCH (one grid, {grid_parts} parts, axis {grid_axis}) seed {seed}: {label} knobs={knobs}: {nd} floats", flush=True)
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
