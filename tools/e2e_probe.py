#!/usr/bin/env python3
"""Developer probe: phases of sweeptt_solve() with pinned host buffers.  usage: e2e_probe.py NSRC"""
import os, sys, pathlib, time, ctypes
os.environ["SWEEPTT_DEBUG_TIMING"] = "1"
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np, torch
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W
nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 14
pageable = len(sys.argv) > 2 and sys.argv[2] == "pageable"
dims = (241, 241, 51)
v = W.heterogeneous_field(dims, 7)
starts = W.starts(111)[:nsrc]
star = P.make_star(W.star("818"))
hv = torch.from_numpy(v).pin_memory()
hout = torch.empty((nsrc,) + dims, dtype=torch.float32)
if not pageable:
    hout = hout.pin_memory()
st_arr = api._make_starts(starts)
ptrs = (ctypes.c_void_p * nsrc)(*[hout[s].data_ptr() for s in range(nsrc)])
opts = api._opts(device=0)
for rep in range(4):
    t0 = time.perf_counter()
    s2 = api.solve_raw(hv.data_ptr(), dims, star, st_arr, ptrs, opts)
    print(f"rep {rep}: wall {1e3*(time.perf_counter()-t0):.2f} ms, solve_ms {s2.solve_ms:.2f}, d2h tail {s2.d2h_ms:.3f}", file=sys.stderr, flush=True)
