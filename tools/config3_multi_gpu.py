#!/usr/bin/env python3
"""BASELINE config 3: 111 sources of the 241x241x51 box sharded over the visible GPUs by ONE call of
sweeptt_solve(num_devices=G) (one host thread + context per GPU, no inter-GPU traffic; the copies of finished
waves run behind the solve of the next ones)."""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W
v = W.heterogeneous_field((241, 241, 51), 7)
starts = W.starts(111)
ndev = P.device_count()
ref = None
for g in [n for n in (1, 2, 4, 8) if n <= ndev]:
    out = torch.empty((111, 241, 241, 51), dtype=torch.float32).pin_memory().numpy()   # pinned: D2H at PCIe speed, overlappable
    P.solve(v, W.star("818"), starts, num_devices=g, out=out)          # warm-up (contexts, graphs)
    t0 = time.perf_counter()
    tt, st = P.solve(v, W.star("818"), starts, num_devices=g, out=out)
    dt = time.perf_counter() - t0
    if ref is None:
        ref = tt.copy()
    same = np.array_equal(ref.view(np.uint32), tt.view(np.uint32))
    print(f"{g} GPU(s): {dt*1e3:8.1f} ms wall (solve {st.solve_ms:7.1f} ms max/device, h2d {st.h2d_ms:.1f}, d2h {st.d2h_ms:.1f}), "
          f"{111/dt:7.1f} sources/s e2e, {st.relaxations/dt/1e9:8.0f} GRelax/s e2e, bit-equal to 1 GPU: {same}", flush=True)
    assert same
