#!/usr/bin/env python3
"""Developer probe: ONE grid over N parts under env knobs.  usage: grid_probe.py NX NY NZ PARTS [K] [KEY=VAL ...]"""
import os, sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
args = [a for a in sys.argv[1:] if "=" not in a]
for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("=", 1); os.environ[k] = v
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W
dims = tuple(int(x) for x in args[:3]); parts = int(args[3]); K = int(args[4]) if len(args) > 4 else None
v = W.heterogeneous_field(dims, 11)
star = P.make_star(W.star("818"))
start = (dims[0] // 2, dims[1] // 2, dims[2] - 1)
best = None
for rep in range(2):
    tt, st = P.solve_slabs(v, star, start, num_slabs=parts, slab_axis=0, rounds_per_poll=K, verbose=1 if rep else 0)
    if best is None or st.solve_ms < best.solve_ms: best = st
print(f"[{' '.join(a for a in sys.argv[1:] if '=' in a)}] {dims} parts={parts} K={K}: solve={best.solve_ms:.1f} ms rounds={best.rounds} tiles={best.tile_visits} "
      f"{best.relaxations/best.solve_ms/1e6:.0f} GRelax/s frac/dev={best.relaxations*4/best.solve_ms/1e9/37.22/best.devices_used:.3f}", flush=True)
