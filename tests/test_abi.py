"""CPU: the C-ABI library loads, exports every symbol include/sweeptt.h declares, and its compute
entry points fail LOUDLY without a CUDA device (no CPU fallback, no oracle behind the product)."""
import ctypes
import pathlib
import re

import numpy as np
import pytest
import torch

import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

ROOT = pathlib.Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "sweeptt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sweeptt_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = P.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sweeptt.h but not exported by libsweeptt.so"
    assert set(syms) == set(api._EXPORTS)


def test_struct_layouts_match_reference():
    assert ctypes.sizeof(P.FS) == 16 and ctypes.sizeof(P.START) == 12  # serial_new/...c:46-58
    assert P.load_library().sweeptt_version().decode().startswith("sweeptt")


def test_product_does_not_route_through_the_oracle():
    """Nothing under the package imports, loads or links the test oracle."""
    for p in (ROOT / "uoparallel_seismic_project_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cpp", ".h", ".inc"}:
            text = p.read_text()
            for needle in ("import oracle", "from oracle", "liboracle", "libref_sweep", "oracle/_ref"):
                assert needle not in text, f"{p} references the oracle ({needle})"
    ldd = __import__("subprocess").run(["ldd", str(P.lib_path())], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd and "libref_sweep" not in ldd


@pytest.mark.skipif(torch.cuda.is_available(), reason="this check is for the GPU-less build container")
def test_compute_fails_loudly_without_a_gpu():
    assert P.device_count() == 0
    v = W.random_field((4, 4, 4))
    with pytest.raises(P.SweepError, match="no CUDA device"):
        P.solve(v, W.star("3"), [(0, 0, 0)])
    with pytest.raises(P.SweepError, match="no CUDA device"):
        P.SweepContext()
    with pytest.raises(P.SweepError, match="no CUDA device"):
        P.solve_slabs(v, W.star("3"), (0, 0, 0), num_slabs=2)
    lib = P.load_library()
    assert not lib.sweeptt_host_alloc(1 << 20), "page-locked memory cannot exist without a CUDA device"
    assert b"no CUDA device" in lib.sweeptt_last_error()
    lib.sweeptt_host_free(None)   # (a null pointer is accepted, like free())


def test_star_distance_helper_matches_reference_formula():
    fs = P.make_star(W.star("818"))
    for l in (0, 100, 817):
        i, j, k = fs[l].i, fs[l].j, fs[l].k
        want = np.float32(10.0) * np.float32(np.sqrt(np.float64(i * i + j * j + k * k)))
        assert np.float32(fs[l].d) == want


def test_header_is_plain_c99_and_links(tmp_path):
    """include/sweeptt.h compiles as strict C99 and a C client links against libsweeptt.so (the binding a reference
    maintainer would add); run without a GPU it must fail loudly, not crash."""
    import subprocess
    src = tmp_path / "client.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "sweeptt.h"
int main(void) {
  struct FS fs[3] = {{1, 0, 0, 0.f}, {0, 1, 0, 0.f}, {0, 0, 1, 0.f}};
  struct START st[1] = {{0, 0, 0}};
  float v[8] = {1, 1, 1, 1, 1, 1, 1, 1}, t[8];
  float *out[1];
  sweeptt_opts opts;
  sweeptt_stats stats;
  int i;
  for (i = 0; i < (int)sizeof opts; i++) ((char *)&opts)[i] = 0;
  opts.struct_size = (int)sizeof opts;
  opts.device = -1;
  out[0] = t;
  sweeptt_star_fill_distances(fs, 3, 10.0f);
  if (fs[0].d != 10.0f) return 3;
  if (sweeptt_device_count() > 0) return 0;  /* (GPU box: covered by the gpu tests) */
  if (sweeptt_solve(v, 2, 2, 2, fs, 3, st, 1, out, &opts, &stats)) return 4;  /* must FAIL without a device */
  printf("%s\n", sweeptt_last_error());
  return 0;
}
''')
    exe = tmp_path / "client"
    lib = P.lib_path()
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{ROOT / 'include'}", "-o", str(exe), str(src),
                    f"-L{lib.parent}", "-lsweeptt", f"-Wl,-rpath,{lib.parent}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    if not torch.cuda.is_available():
        assert "no CUDA device" in r.stdout
