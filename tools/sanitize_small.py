#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck): tiled + simple kernels, slabs, groups."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W
v = W.heterogeneous_field((19, 18, 37), seed=2)
for star in ("818", "3"):
    for kernel in (api.KERNEL_TILED, api.KERNEL_SIMPLE):
        tt, st = P.solve(v, W.star(star), [(9, 9, 36), (0, 0, 0), (18, 17, 0)], kernel=kernel)
        print(star, kernel, st.rounds, float(tt[0].max()))
tt, st = P.solve_slabs(v, W.star("818"), (9, 9, 36), num_slabs=3, slab_axis=0)
print("slabs", st.rounds, float(tt.max()))
