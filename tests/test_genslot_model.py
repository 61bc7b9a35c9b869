"""CPU: randomised interleaving model of the single-launch scheduler's work-list slots (kernels.cu SolveState::gslot).
With the shipped protocol -- entry count and pop cursor in ONE 64-bit word, popped with one fetch_add, recycled with
one exchange -- every entry of every generation is handed out exactly once, even to a CTA that is still four
generations behind; round 1's protocol (separate words, two stores) is caught handing entries out twice."""
import re
import subprocess

from conftest import ROOT


def _run(exe, protocol, runs):
    r = subprocess.run([str(exe), protocol, str(runs)], capture_output=True, text=True, timeout=600)
    m = re.search(r"double_handouts (\d+) never_handed_out (\d+)", r.stdout)
    return int(m.group(1)), int(m.group(2))


def test_slot_recycling_protocol(tmp_path):
    exe = tmp_path / "genslot_model"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "native" / "genslot_model.cpp")], check=True)
    assert _run(exe, "word", 1500) == (0, 0)
    double, lost = _run(exe, "split", 1500)
    assert double > 0 and lost == 0, "the model should reproduce round 1's double hand-out"
