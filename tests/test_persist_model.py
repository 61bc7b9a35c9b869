"""CPU: randomised interleaving model of the single-launch scheduler's termination protocol (kernels.cu
build_generation + the finisher's epilogue): with the kernel's ordering -- in-flight count read BEFORE the key scan,
busy tiles left pending, neighbours' keys set before busy is cleared and before the count drops -- `done` is never
announced while a key is pending or a tile in flight, no tile is relaxed by two CTAs at once and every run ends; reading
the in-flight count after the scan instead is caught announcing `done` too early."""
import re
import subprocess

from conftest import ROOT


def _run(exe, mode, runs):
    r = subprocess.run([str(exe), mode, str(runs)], capture_output=True, text=True, timeout=600)
    m = re.search(r"early_done (\d+) overlaps (\d+) hung (\d+)", r.stdout)
    return tuple(int(x) for x in m.groups())


def test_termination_protocol_of_the_single_launch_scheduler(tmp_path):
    exe = tmp_path / "persist_model"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "native" / "persist_model.cpp")], check=True)
    assert _run(exe, "shipped", 3000) == (0, 0, 0)
    early, overlaps, hung = _run(exe, "late", 3000)
    assert early > 0 and overlaps == 0 and hung == 0, "the model should catch the in-flight count being read after the scan"
