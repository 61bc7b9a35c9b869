import sys, time, ctypes, pathlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W
DIMS=(241,241,51)
v=W.heterogeneous_field(DIMS,7); starts=W.starts(4); star=P.make_star(W.star("818"))
hv=torch.from_numpy(v).pin_memory(); hout=torch.empty((4,)+DIMS,dtype=torch.float32).pin_memory()
st_arr=api._make_starts(starts); ptrs=(ctypes.c_void_p*4)(*[hout[s].data_ptr() for s in range(4)]); opts=api._opts(device=0)
for _ in range(3): api.solve_raw(hv.data_ptr(), DIMS, star, st_arr, ptrs, opts)
for _ in range(5):
    t0=time.perf_counter(); s=api.solve_raw(hv.data_ptr(), DIMS, star, st_arr, ptrs, opts); dt=(time.perf_counter()-t0)*1e3
    print(f"wall {dt:.2f} ms: h2d {s.h2d_ms:.2f} solve {s.solve_ms:.2f} (kernel {s.relax_kernel_ms:.2f}) d2h {s.d2h_ms:.2f}")
