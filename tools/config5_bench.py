#!/usr/bin/env python3
"""BASELINE config 5: ONE source on ONE huge grid spread over the GPUs of the box (sweeptt_solve_slabs: shared box in
peer memory, block-cyclic ownership, halo reads fused into the relaxation kernel's TMA loads).

  python tools/config5_bench.py [--dims 2401 2401 501] [--parts 1 2 4 8] [--reps 2] [--seed 13]

ONE process drives all devices (not torchrun).  Prints one JSON line per part count: solve time (host wall clock
between "all devices ready" and "quiescent"), executed GRelax/s, speed-up against the first entry, and whether the
field is bit-identical to the first entry's (sha256) -- strong scaling of the mpi/16partsmpi.c workload."""
import argparse
import hashlib
import json
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import uoparallel_seismic_project_b200 as P  # noqa: E402
from uoparallel_seismic_project_b200 import workloads as W  # noqa: E402


def big_field(dims, seed):
    """workloads.heterogeneous_field without the full-size temporaries (11.5 GB per copy at 2401x2401x501)."""
    nx, ny, nz = dims
    rng = np.random.default_rng(seed)
    grad = np.linspace(0.30, 0.15, nz, dtype=np.float32)
    v = rng.random(dims, dtype=np.float32)
    v *= np.float32(0.2)
    v += np.float32(0.9)
    v *= grad[None, None, :]
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, nargs=3, default=[2401, 2401, 501])
    ap.add_argument("--parts", type=int, nargs="+", default=None)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--seed", type=int, default=13)
    ap.add_argument("--axis", type=int, default=0)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "config5.jsonl"))
    a = ap.parse_args()
    dims = tuple(a.dims)
    ndev = P.device_count()
    parts = a.parts or [n for n in (1, 2, 4, 8) if n <= ndev]
    t0 = time.time()
    v = big_field(dims, a.seed)
    assert v.dtype == np.float32
    small = W.heterogeneous_field((16, 16, 16), a.seed)
    assert np.array_equal(big_field((16, 16, 16), a.seed), small), "big_field must equal workloads.heterogeneous_field"
    print(f"model {dims} generated in {time.time() - t0:.1f} s", flush=True)
    start = (dims[0] // 2, dims[1] // 2, dims[2] - 1)      # centre of the bottom face
    star = P.make_star(W.star("818"))
    first = None
    pathlib.Path(a.out).parent.mkdir(exist_ok=True)
    with open(a.out, "a") as log:
        for n in parts:
            best = None
            for rep in range(a.reps):
                tt, st = P.solve_slabs(v, star, start, num_slabs=n, slab_axis=a.axis)
                if best is None or st.solve_ms < best.solve_ms:
                    best = st
            sha = hashlib.sha256(tt.tobytes()).hexdigest()
            if first is None:
                first = (best.solve_ms, sha, n)
            line = {"workload": f"{dims[0]}x{dims[1]}x{dims[2]} heterogeneous (seed {a.seed}), 818-FS, 1 source at {start}",
                    "parts": n, "devices_used": best.devices_used, "solve_ms": best.solve_ms,
                    "setup_upload_ms": best.h2d_ms, "gather_ms": best.d2h_ms, "rounds_max": best.rounds,
                    "tile_visits": best.tile_visits, "grelax": best.relaxations / 1e9,
                    "grelax_per_s": best.relaxations / best.solve_ms / 1e6,
                    "frac_of_fp32_issue_roof_per_device": best.relaxations * 4 / best.solve_ms / 1e9 / 37.22 / max(1, best.devices_used),
                    "speedup_vs_first": first[0] / best.solve_ms, "first_parts": first[2],
                    "sha256": sha, "bit_equal_to_first": sha == first[1]}
            print(json.dumps(line), flush=True)
            log.write(json.dumps(line) + "\n")
            del tt


if __name__ == "__main__":
    main()
