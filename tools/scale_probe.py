#!/usr/bin/env python3
"""Kernel efficiency at scale: one source on a 601x601x126 box (many tiles per round)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W
dims = (601, 601, 126)
v = W.heterogeneous_field(dims, seed=13)
with P.SweepContext() as ctx:
    ctx.set_model(v); ctx.set_star(W.star("818")); ctx.set_sources([(300, 300, 125), (10, 10, 0), (590, 300, 60), (300, 590, 125)])
    for _ in range(2):
        st = ctx.run()
    print(f"{st.solve_ms:.1f} ms, rounds {st.rounds}, {st.relaxations/1e9:.0f} GRelax -> {st.relaxations/st.solve_ms/1e6:.0f} GRelax/s")
