#!/usr/bin/env python3
"""Developer probe: device-resident solve of the 241x241x51 box under env knobs.
Usage: probe.py NSRC [STAR] [KEY=VAL ...]   (NSRC 4 = start-4, else the first NSRC rows of start-111)"""
import os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
args = [a for a in sys.argv[1:] if "=" not in a]
for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("=", 1)
        os.environ[k] = v
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

nsrc = int(args[0]) if args else 4
star = args[1] if len(args) > 1 else "818"
v = W.constant_field((241, 241, 51)) if os.environ.get("PROBE_CONST") else W.heterogeneous_field((241, 241, 51), 7)
starts = W.starts(4) if nsrc == 4 else W.starts(111)[:nsrc]
with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
    ctx.set_model(v); ctx.set_star(W.star(star)); ctx.set_sources(starts)
    best = None
    for rep in range(6):
        st = ctx.run()
        if rep >= 2 and (best is None or st.solve_ms < best.solve_ms):
            best = st
    st = best
    full = ctx.relaxations_per_round * nsrc
    viol = sum(ctx.count_violations(s) for s in range(min(nsrc, 4)))
    knobs = " ".join(a for a in sys.argv[1:] if "=" in a)
    print(f"[{knobs}] nsrc={nsrc} star={star}: solve={st.solve_ms:.2f}ms launches={st.relax_launches} "
          f"({st.relaxations/full:.2f} grid-equivalents) tiles={st.tile_visits} -> {st.relaxations/st.solve_ms/1e6:.1f} GRelax/s, "
          f"{nsrc/st.solve_ms*1e3:.1f} sources/s, frac={st.relaxations*4/st.solve_ms/1e9/37.22:.3f} units {st.units_changed}/{st.units_run} viol={viol}",
          flush=True)
