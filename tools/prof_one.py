#!/usr/bin/env python3
"""Short run for ncu: one source, batched loop, N rounds."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W
nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 30
star = sys.argv[3] if len(sys.argv) > 3 else "818"
v = W.heterogeneous_field((241, 241, 51), 7)
with P.SweepContext(kernel=api.KERNEL_TILED, loop=api.LOOP_BATCHED) as ctx:
    ctx.set_model(v); ctx.set_star(W.star(star)); ctx.set_sources(W.starts(111)[:nsrc])
    ctx.reset()
    ch, st = ctx.step(rounds)
    print("rounds", st.rounds, "ms", st.solve_ms, "GRelax", st.relaxations / 1e9, "tiles", st.tile_visits)
