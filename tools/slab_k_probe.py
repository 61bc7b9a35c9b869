#!/usr/bin/env python3
"""Developer probe: rounds between halo exchanges (rounds_per_poll) of the slab decomposition.
usage: slab_k_probe.py SCALE "8,16,32" """
import sys, time, hashlib, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
ks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8,16,32").split(",")]
dims = tuple(int(round(d * scale)) for d in (2401, 2401, 501))
start = (dims[0] // 2, dims[1] // 2, dims[2] - 1)
v = W.heterogeneous_field(dims, 13)
off = W.star("818")
one, st1 = P.solve(v, off, [start])
print(f"{dims} 1 GPU: {st1.solve_ms:.0f} ms", flush=True)
h1 = hashlib.sha256(one[0].tobytes()).hexdigest()
P.load_library().sweeptt_release_cache()
for g in [n for n in (2, 4, 8) if n <= P.device_count()]:
    for k in ks:
        tt, st = P.solve_slabs(v, off, start, num_slabs=g, slab_axis=0, rounds_per_poll=k)
        same = hashlib.sha256(tt.tobytes()).hexdigest() == h1
        print(f"{g} GPUs K={k}: {st.solve_ms:.0f} ms, {st.relaxations/1e12:.2f} TRelax, rounds {st.rounds}, bit-equal {same}", flush=True)
