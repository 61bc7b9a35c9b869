#!/usr/bin/env python3
"""End-to-end time of the drop-in command on BASELINE config 3: load .vbox -> solve 111 sources -> write output.tt
(serial_new/sweep-tt-multistart.c:12 command line; :176-194 output).  Writes gpurun_out/cli_e2e.json."""
import json
import os
import pathlib
import re
import subprocess
import sys
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import uoparallel_seismic_project_b200 as P  # noqa: E402
from uoparallel_seismic_project_b200 import workloads as W  # noqa: E402

nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 111
exe = ROOT / "uoparallel_seismic_project_b200" / "lib" / "sweep-tt-multistart"
v = W.heterogeneous_field((241, 241, 51), 7)
with tempfile.TemporaryDirectory(dir=os.environ.get("TMPDIR", "/tmp")) as td:
    td = pathlib.Path(td)
    P.vbox_store(td / "m.vbox", v, origin=(1, 1, 1))
    W.write_star_file(td / "818-FS.txt", W.star("818"))
    W.write_start_file(td / "start.txt", W.starts(111)[:nsrc])
    res = []
    for rep in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([str(exe), "m.vbox", "818-FS.txt", "start.txt"], cwd=td, capture_output=True, text=True,
                           env=dict(os.environ, SWEEPTT_TIMING="1"), timeout=1200)
        wall = time.perf_counter() - t0
        m = re.search(r"solve_call_s=([\d.]+) output_tt_s=([\d.]+)", r.stderr)
        size = (td / "output.tt").stat().st_size
        res.append({"wall_s": wall, "solve_call_s": float(m.group(1)), "output_tt_s": float(m.group(2)),
                    "output_tt_bytes": size, "output_tt_MB_per_s": size / 1e6 / float(m.group(2)), "returncode": r.returncode,
                    "load_and_startup_s": wall - float(m.group(1)) - float(m.group(2))})
        print(res[-1], flush=True)
out = {"workload": f"config 3 CLI: 241x241x51 .vbox, 818-FS, {nsrc} sources -> output.tt", "runs": res}
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "cli_e2e.json").write_text(json.dumps(out, indent=1))
