#!/usr/bin/env python3
"""Probe: does splitting the sources of config 2 over several concurrently running contexts
(own stream + CUDA graph each) hide the per-round tail?"""
import sys, time, pathlib, threading
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W
v = W.heterogeneous_field((241, 241, 51), 7)
starts = W.starts(4)
for groups in (1, 2, 4):
    ctxs = []
    for g in range(groups):
        c = P.SweepContext()
        c.set_model(v); c.set_star(W.star("818")); c.set_sources(starts[g::groups])
        ctxs.append(c)
    def run_all():
        th = [threading.Thread(target=c.run) for c in ctxs]
        [t.start() for t in th]; [t.join() for t in th]
    for _ in range(3): run_all()
    t0 = time.perf_counter()
    for _ in range(10): run_all()
    dt = (time.perf_counter() - t0) / 10
    print(f"groups={groups}: {dt*1e3:.2f} ms per 4-source solve -> {4/dt:.1f} sources/s", flush=True)
    [c.close() for c in ctxs]
