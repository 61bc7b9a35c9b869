#!/usr/bin/env python3
"""Pack the reference's input fixtures (forward stars, start points) into data/inputs.npz.

Run ONCE in the build container (it reads /root/reference/docs, which does not exist on
the GPU box).  The arrays are data the drop-in surface consumes (SURVEY.md §2 row 12:
docs/{3,5,818}-FS.txt, docs/start-*-241-241-51.txt); they are stored as int32 arrays and
re-materialised as text files in the reference's formats by
uoparallel_seismic_project_b200.workloads.write_star_file / write_start_file.
"""
import pathlib
import sys

import numpy as np

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/docs")
OUT = pathlib.Path(__file__).resolve().parents[1] / "data" / "inputs.npz"


def read_counted(path):
    toks = path.read_text().split()
    n = int(toks[0])
    a = np.array(toks[1:1 + 3 * n], dtype=np.int32).reshape(n, 3)
    return a


arrays = {}
for name in ("3-FS", "5-FS", "818-FS"):
    arrays["fs_" + name.split("-")[0]] = read_counted(REF / f"{name}.txt")
for n in (1, 4, 10, 24, 111):
    arrays[f"start_{n}"] = read_counted(REF / f"start-{n}-241-241-51.txt")
np.savez_compressed(OUT, **arrays)
for k, v in arrays.items():
    print(k, v.shape)
