#!/usr/bin/env python3
"""Race hunt for the single-launch scheduler: the same solves many times under different early-build
distances and bucket factors; every result must be bit-identical to the first and a fixed point."""
import hashlib, os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cases = [((241, 241, 51), "818", list(map(tuple, W.starts(4)))),
         ((97, 83, 61), "818", [(48, 41, 60), (0, 0, 0), (96, 82, 30)]),
         ((150, 40, 33), "5", [(0, 0, 0), (149, 39, 32)]),
         ((64, 64, 64), "3", [(31, 31, 31)])]
bad = 0
for dims, star, starts in cases:
    v = W.heterogeneous_field(dims, 7)
    ref = None
    for r in range(reps):
        os.environ["SWEEPTT_PERSIST"] = "1"  # (the 3-FS kernel defaults to the graph of rounds)
        os.environ["SWEEPTT_LOOKAHEAD"] = ["8", "0", "0.05", "2", "30"][r % 5]
        os.environ["SWEEPTT_BUCKET"] = ["2", "0.5", "5", "-1"][r % 4]
        os.environ["SWEEPTT_TRIGGER_FRAC"] = ["0.4", "0", "0.9"][r % 3]
        with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
            ctx.set_model(v); ctx.set_star(W.star(star)); ctx.set_sources(starts)
            st = ctx.run()
            assert st.relax_launches == 1, "not the single-launch path"
            h = hashlib.sha256(b"".join(ctx.get_tt(s).tobytes() for s in range(len(starts)))).hexdigest()
            viol = sum(ctx.count_violations(s) for s in range(len(starts)))
        ref = ref or h
        if h != ref or viol:
            bad += 1
            print(f"MISMATCH {dims} {star}-FS rep {r}: violations {viol}", flush=True)
    print(f"{dims} {star}-FS: {reps} solves, sha {ref[:16]}", flush=True)
print("bad =", bad)
sys.exit(1 if bad else 0)
