#!/usr/bin/env python3
"""BASELINE configs 4 and 5 on one B200 (device-resident; fixed-point verifier instead of the oracle).
usage: config45.py 4|5 [nsources]"""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "4"
if which == "4":
    dims, seed = (1201, 1201, 251), 11
    nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    starts = W.starts(24)[:nsrc] * np.array([5, 5, 5]) // 1   # start-24 coordinates scaled x5 (k = 250)
else:
    dims, seed = (2401, 2401, 501), 13
    nsrc = 1
    starts = np.array([[1200, 1200, 500]])
t0 = time.time()
v = W.heterogeneous_field(dims, seed)
print(f"field {dims} generated in {time.time()-t0:.1f}s", flush=True)
with P.SweepContext() as ctx:
    t0 = time.time()
    ctx.set_model(v); ctx.set_star(W.star("818")); ctx.set_sources(starts)
    print(f"setup {time.time()-t0:.1f}s, pool {ctx.pool_bytes/2**30:.1f} GiB, {ctx.relaxations_per_round/1e9:.1f} GRelax per full round per source", flush=True)
    for rep in range(2):
        st = ctx.run()
        print(f"run {rep}: {st.solve_ms:.1f} ms, rounds {st.rounds}, {st.relaxations/1e12:.2f} TRelax "
              f"({st.relaxations/ctx.relaxations_per_round/nsrc:.1f} full rounds/source) -> {st.relaxations/st.solve_ms/1e6:.0f} GRelax/s, "
              f"{nsrc/st.solve_ms*1e3:.2f} sources/s", flush=True)
    t0 = time.time()
    viol = [ctx.count_violations(s) for s in range(min(nsrc, 4))]
    print("fixed-point violations (first sources):", viol, f"{time.time()-t0:.1f}s")
    assert all(x == 0 for x in viol)
