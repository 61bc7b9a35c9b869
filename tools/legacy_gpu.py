#!/usr/bin/env python3
"""Time the REFERENCE'S OWN CUDA programs on this box's GPU ("the old GPU code, same B200").

oracle/_ref/legacy_cudasweep_{230,380} are cuda/cudasweep-tt-multistart_230.cu / _380.cu compiled unmodified
(oracle/Makefile `legacy`).  They read text dialect A, want exactly 4 start points (STARTMAX 4) and print one line
per sweep: " start point: s, sweep n: c changes, sweep <kernel ms>, data trans <D2H ms>".  _230 is the fastest
single-GPU variant but drops the start-skip (not parity exact); _380 keeps the serial edge set.
Both update travel times in place without atomics, so their sweep counts vary from run to run.

Writes gpurun_out/legacy_gpu.json (copy it to profiles/ to have bench.py cite it as `legacy_gpu`).
Measurement only: nothing here is on the product path.
"""
import hashlib
import json
import pathlib
import re
import subprocess
import sys
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import uoparallel_seismic_project_b200 as P  # noqa: E402
from uoparallel_seismic_project_b200 import workloads as W  # noqa: E402

LINE = re.compile(r"start point: (\d+), sweep (\d+): (-?\d+) changes, sweep ([\d.]+), data trans ([\d.]+)")


def main():
    dims = (241, 241, 51)
    v = W.heterogeneous_field(dims, 7)
    starts = W.starts(4)
    out = {"workload": "config 2: 241x241x51 heterogeneous (seed 7), 818-FS, start-4", "variants": {}}
    with tempfile.TemporaryDirectory() as td:
        td = pathlib.Path(td)
        t0 = time.time()
        # text dialect A, written with numpy (2.96 M lines)
        x, y, z = np.meshgrid(np.arange(1, dims[0] + 1), np.arange(1, dims[1] + 1), np.arange(1, dims[2] + 1), indexing="ij")
        np.savetxt(td / "v.txt", np.column_stack([x.ravel(), y.ravel(), z.ravel(), v.ravel().astype(np.float64)]),
                   fmt=["%d", "%d", "%d", "%.9g"], delimiter=",")
        W.write_star_file(td / "818-FS.txt", W.star("818"))
        W.write_start_file(td / "start-4.txt", starts)
        print(f"inputs written in {time.time() - t0:.1f} s", flush=True)
        # our own converged fields in the same text format, for a byte comparison of output.tt
        tt, st = P.solve(v, W.star("818"), starts)
        P.write_output_tt(td / "ours.tt", tt)
        ours_sha = hashlib.sha256((td / "ours.tt").read_bytes()).hexdigest()
        out["ours"] = {"solve_ms_4_sources": st.solve_ms, "output_tt_sha256": ours_sha}
        for name in ("230", "380"):
            exe = ROOT / "oracle" / "_ref" / f"legacy_cudasweep_{name}"
            if not exe.exists():
                out["variants"][name] = {"error": "binary not built"}
                continue
            t0 = time.time()
            try:
                r = subprocess.run([str(exe), "v.txt", "818-FS.txt", "start-4.txt"], cwd=td, capture_output=True,
                                   text=True, timeout=900)
            except subprocess.TimeoutExpired:
                out["variants"][name] = {"error": "timeout after 900 s"}
                continue
            wall = time.time() - t0
            per = {}
            for m in LINE.finditer(r.stdout):
                s = int(m.group(1))
                per.setdefault(s, []).append((float(m.group(4)), float(m.group(5))))
            if not per:
                out["variants"][name] = {"error": "no sweep lines", "stdout_tail": r.stdout[-400:], "stderr_tail": r.stderr[-400:]}
                continue
            sweeps = [len(per[s]) for s in sorted(per)]
            k_ms = [sum(a for a, _ in per[s]) for s in sorted(per)]
            d_ms = [sum(b for _, b in per[s]) for s in sorted(per)]
            allk = [a for s in per for a, _ in per[s]]
            res = {"sweeps_per_source": sweeps, "kernel_ms_per_sweep_mean": sum(allk) / len(allk),
                   "kernel_ms_per_sweep_min": min(allk), "kernel_ms_per_source": k_ms, "d2h_flag_ms_per_source": d_ms,
                   "ms_per_converged_source_mean": (sum(k_ms) + sum(d_ms)) / len(k_ms),
                   "grelax_per_s_kernel": 2_246_171_812 / (sum(allk) / len(allk) * 1e-3) / 1e9,
                   "process_wall_s": wall, "returncode": r.returncode}
            tt_file = td / "output.tt"
            if tt_file.exists():
                res["output_tt_sha256"] = hashlib.sha256(tt_file.read_bytes()).hexdigest()
                res["output_tt_equals_ours"] = res["output_tt_sha256"] == ours_sha
                if not res["output_tt_equals_ours"]:
                    a = (td / "ours.tt").read_text().splitlines()
                    b = tt_file.read_text().splitlines()
                    res["output_tt_lines_differing"] = sum(1 for p, q in zip(a, b) if p != q) + abs(len(a) - len(b))
                tt_file.unlink()
            out["variants"][name] = res
            print(name, json.dumps(res)[:600], flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "legacy_gpu.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
