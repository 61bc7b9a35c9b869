#!/usr/bin/env python3
"""Developer timing probe (not the bench contract): config-2-like solve with per-kernel timing."""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

star = sys.argv[1] if len(sys.argv) > 1 else "818"
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kind = sys.argv[3] if len(sys.argv) > 3 else "hetero"
v = W.heterogeneous_field((241, 241, 51), 7) if kind == "hetero" else W.constant_field((241, 241, 51))
starts = W.starts(4) if nsrc == 4 else W.starts(111)[:nsrc]
for loop, prof in ((api.LOOP_BATCHED, 1), (api.LOOP_BATCHED, 0), (api.LOOP_GRAPH, 0)):
    with P.SweepContext(kernel=api.KERNEL_TILED, loop=loop, profile_kernels=prof) as ctx:
        ctx.set_model(v); ctx.set_star(W.star(star)); ctx.set_sources(starts)
        for rep in range(3):
            st = ctx.run()
        full = ctx.relaxations_per_round * nsrc
        print(f"loop={loop} prof={prof}: rounds={st.rounds} solve={st.solve_ms:.2f}ms relax_kernel={st.relax_kernel_ms:.2f}ms "
              f"GRelax={st.relaxations/1e9:.2f} ({st.relaxations/full:.1f} full rounds) tiles={st.tile_visits} "
              f"-> {st.relaxations/st.solve_ms/1e6:.1f} GRelax/s, {nsrc/st.solve_ms*1e3:.1f} sources/s "
              f"units changed/run = {st.units_changed}/{st.units_run}", flush=True)
        if prof:
            print("   violations:", [ctx.count_violations(s) for s in range(nsrc)])
