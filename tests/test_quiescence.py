"""CPU: the barrier-free termination detector of the multi-device grid solver (csrc/quiescence.h) against a randomised
model of parts that run batches at different speeds and wake each other up -- it must never announce the end while a
wake-up is still unprocessed, and must announce it once everything is quiet.  The model is sharp enough to reject the
detector when it asks for fewer than two further batches per part."""
import re
import subprocess

from conftest import ROOT

CSRC = ROOT / "uoparallel_seismic_project_b200" / "csrc"
MODEL = ROOT / "tests" / "native" / "quiescence_model.cpp"


def _run(tmp_path, header_dir, runs):
    exe = tmp_path / f"model_{header_dir.name}"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{header_dir}", "-o", str(exe), str(MODEL), "-lpthread"], check=True)
    r = subprocess.run([str(exe), str(runs)], capture_output=True, text=True, timeout=600)
    m = re.search(r"false_ends (\d+) never_ended (\d+)", r.stdout)
    return int(m.group(1)), int(m.group(2))


def test_termination_detector_never_ends_early_and_always_ends(tmp_path):
    assert _run(tmp_path, CSRC, 5000) == (0, 0)


def test_the_model_rejects_a_detector_that_waits_for_fewer_batches(tmp_path):
    text = (CSRC / "quiescence.h").read_text()
    assert "seq[q] < seq0[q] + 2" in text
    for fewer in (0, 1):
        d = tmp_path / f"mut{fewer}"
        d.mkdir()
        (d / "quiescence.h").write_text(text.replace("seq[q] < seq0[q] + 2", f"seq[q] < seq0[q] + {fewer}"))
        false_ends, never = _run(tmp_path, d, 5000)
        assert false_ends > 0 and never == 0, f"+{fewer} batches should end early in the model"
