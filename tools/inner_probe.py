#!/usr/bin/env python3
"""Developer probe: in-tile passes per visit (SWEEPTT_INNER) for the small stars."""
import os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W
for star, kind, nsrc in (("3", "const", 1), ("3", "hetero", 4), ("5", "hetero", 4)):
    v = W.heterogeneous_field((241, 241, 51), 7) if kind == "hetero" else W.constant_field((241, 241, 51))
    starts = W.starts(4) if nsrc == 4 else W.starts(111)[:nsrc]
    for inner in ("1", "2", "3", "4"):
        for persist in ("1", "0"):
            os.environ["SWEEPTT_INNER"] = inner
            os.environ["SWEEPTT_PERSIST"] = persist
            with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
                ctx.set_model(v); ctx.set_star(W.star(star)); ctx.set_sources(starts)
                best = min((ctx.run() for _ in range(5)), key=lambda s: s.solve_ms)
                print(f"{star}-FS {kind} {nsrc} src inner={inner} persist={persist}: {best.solve_ms:.2f} ms, "
                      f"{best.relaxations/ctx.relaxations_per_round/nsrc:.1f} full rounds, tiles {best.tile_visits}", flush=True)
