#!/usr/bin/env python3
"""Developer probe: config-2-like solve under combinations of the tuning knobs
(SWEEPTT_GROUPS x SWEEPTT_BUCKET [x extra env]).  Usage: knob_probe.py NSRC "1,2,4" "1,2,3" [KEY=VAL ...]"""
import os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
groups = sys.argv[2].split(",") if len(sys.argv) > 2 else ["2"]
buckets = sys.argv[3].split(",") if len(sys.argv) > 3 else ["2"]
for kv in sys.argv[4:]:
    k, v = kv.split("=", 1)
    os.environ[k] = v
v = W.heterogeneous_field((241, 241, 51), 7)
starts = W.starts(4) if nsrc == 4 else W.starts(111)[:nsrc]
for g in groups:
    for b in buckets:
        os.environ["SWEEPTT_GROUPS"] = g
        os.environ["SWEEPTT_BUCKET"] = b
        with P.SweepContext(kernel=api.KERNEL_TILED, loop=api.LOOP_GRAPH) as ctx:
            ctx.set_model(v); ctx.set_star(W.star("818")); ctx.set_sources(starts)
            best = None
            for rep in range(6):
                st = ctx.run()
                if rep >= 2 and (best is None or st.solve_ms < best.solve_ms):
                    best = st
            st = best
            full = ctx.relaxations_per_round * nsrc
            print(f"groups={g} bucket={b}: rounds={st.rounds} solve={st.solve_ms:.2f}ms ({st.relaxations/full:.2f} full rounds) "
                  f"tiles={st.tile_visits} -> {st.relaxations/st.solve_ms/1e6:.1f} GRelax/s, {nsrc/st.solve_ms*1e3:.1f} sources/s "
                  f"units {st.units_changed}/{st.units_run}", flush=True)
