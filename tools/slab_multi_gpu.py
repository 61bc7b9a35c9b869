#!/usr/bin/env python3
"""Slab decomposition across REAL devices: bit-equality with the single-GPU field + timings.
usage: slab_multi_gpu.py [nx ny nz] (default 601 601 126 = config 5 scaled by 1/4 per axis)"""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W

dims = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (601, 601, 126)
ndev = P.device_count()
v = W.heterogeneous_field(dims, seed=13)
start = (dims[0] // 2, dims[1] // 2, dims[2] - 1)
off = W.star("818")
t0 = time.time()
one, st1 = P.solve(v, off, [start])
print(f"1 GPU : {st1.solve_ms:9.1f} ms device, {time.time()-t0:6.2f} s wall, rounds {st1.rounds}, "
      f"{st1.relaxations/1e9:8.1f} GRelax -> {st1.relaxations/st1.solve_ms/1e6:7.1f} GRelax/s", flush=True)
for g in [n for n in (2, 4, 8) if n <= ndev]:
    for axis in (0, 2):
        t0 = time.time()
        tt, st = P.solve_slabs(v, off, start, num_slabs=g, slab_axis=axis)
        same = np.array_equal(tt.view(np.uint32), one[0].view(np.uint32))
        print(f"{g} GPUs axis {axis}: {st.solve_ms:9.1f} ms solve, {time.time()-t0:6.2f} s wall, {st.relaxations/1e9:8.1f} GRelax "
              f"-> {st.relaxations/st.solve_ms/1e6:7.1f} GRelax/s, bit-equal to 1 GPU: {same}", flush=True)
        assert same
