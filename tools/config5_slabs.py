#!/usr/bin/env python3
"""BASELINE config 5 at full size: 2401x2401x501, one source, slab decomposition over the visible GPUs.
Correctness: device fixed-point verifier is not available across slabs, so the gathered field is compared
with the single-GPU field by sha256 (which tools/config45.py verified to be a fixed point)."""
import sys, time, hashlib, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dims = tuple(int(round(d * scale)) for d in (2401, 2401, 501))
start = (dims[0] // 2, dims[1] // 2, dims[2] - 1)
t0 = time.time(); v = W.heterogeneous_field(dims, 13); print(f"field {dims} in {time.time()-t0:.1f}s", flush=True)
off = W.star("818")
t0 = time.time(); one, st1 = P.solve(v, off, [start]); 
print(f"1 GPU: solve {st1.solve_ms:.0f} ms ({st1.relaxations/st1.solve_ms/1e6:.0f} GRelax/s), wall {time.time()-t0:.1f}s", flush=True)
h1 = hashlib.sha256(one[0].tobytes()).hexdigest()
P.load_library().sweeptt_release_cache()
for g in [n for n in (2, 4, 8) if n <= P.device_count()]:
    t0 = time.time()
    tt, st = P.solve_slabs(v, off, start, num_slabs=g, slab_axis=0)
    same = hashlib.sha256(tt.tobytes()).hexdigest() == h1
    print(f"{g} GPUs (x slabs): solve {st.solve_ms:.0f} ms ({st.relaxations/st.solve_ms/1e6:.0f} GRelax/s aggregate), "
          f"wall {time.time()-t0:.1f}s, bit-equal to 1 GPU: {same}", flush=True)
    assert same
