"""CPU: the offline model of the tile scheduler behind DESIGN.md 4b (tools/sim/worksim.c) builds, converges to the
same field with and without its column-skipping criterion (the criterion is exact), and reports the fractions the
design decision rests on."""
import re
import subprocess

import numpy as np

from uoparallel_seismic_project_b200 import workloads as W

from conftest import ROOT


def test_work_reduction_model_is_exact_and_reports_its_fractions(tmp_path):
    exe = tmp_path / "worksim"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tools" / "sim" / "worksim.c"), "-lm"],
                   check=True)
    dims = (20, 41, 43)
    v = W.heterogeneous_field(dims, 7)
    v.tofile(tmp_path / "v.f32")
    W.write_star_file(tmp_path / "star.txt", W.star("818"))
    r = subprocess.run([str(exe), str(tmp_path / "v.f32"), *map(str, dims), str(tmp_path / "star.txt"), "19", "20", "21"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-400:] + r.stderr[-400:]
    assert "fields differ in 0 of" in r.stdout
    m = re.search(r"columns needed ([\d.]+), single pulls needed ([\d.]+); neighbour-tile mask: columns needed ([\d.]+)", r.stdout)
    exact_cols, exact_pulls, mask_cols = map(float, m.groups())
    assert exact_pulls < exact_cols <= mask_cols <= 1.0       # coarser criteria keep more work
    runs = re.findall(r"skip=(\d): generations \d+, tile visits (\d+)", r.stdout)
    assert runs[0][1] == runs[1][1]                           # skipping columns never changes the schedule
