import json
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the C-ABI library and the oracle exist (build() is idempotent and quick)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden_small():
    """Fields produced by the reference's own serial_new code (tools/make_golden.py small)."""
    from uoparallel_seismic_project_b200 import workloads as W
    z = np.load(ROOT / "tests" / "golden" / "small_cases.npz")
    meta = json.loads(bytes(z["meta_json"]).decode())
    cases = []
    for m in meta:
        m = dict(m)
        m["tt"] = z[f"{m['case']}__{m['idx']}"]
        m["v_sha"] = bytes(z[f"{m['case']}__v_sha"])
        cases.append(m)
    return cases


def make_field(kind, dims, seed):
    from uoparallel_seismic_project_b200 import workloads as W
    dims = tuple(dims)
    return {"random": lambda: W.random_field(dims, seed), "constant": lambda: W.constant_field(dims, 0.25),
            "hetero": lambda: W.heterogeneous_field(dims, seed), "contrast": lambda: W.contrast_field(dims, seed)}[kind]()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = bits(a), bits(b)
    assert a.shape == b.shape, what
    nd = int((a != b).sum())
    assert nd == 0, f"{what}: {nd} of {a.size} floats differ"
