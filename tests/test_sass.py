"""CPU: inspect the SASS of the shipped library -- evidence that the hot kernel is the
Blackwell-native one (TMA) and that nothing in it can break the bit-exactness contract
(no scalar FFMA; every packed FFMA2 only adds the broadcast run-time scalar -0.0, never a packed pair)."""
import re
import shutil
import subprocess

import pytest

import uoparallel_seismic_project_b200 as P

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


def _sass():
    out = subprocess.run(["cuobjdump", "-sass", str(P.lib_path())], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            funcs[name].append(line)
    return funcs


def test_relax_kernels_are_tma_and_never_fuse_multiply_add():
    funcs = _sass()
    relax = {k: v for k, v in funcs.items() if "relax_tiled" in k}
    assert len(relax) >= 6  # 3 generic halo variants + 3 stock stars
    for name, lines in relax.items():
        text = "\n".join(lines)
        assert "UTMALDG" in text, f"{name}: no TMA load"
        assert "SYNCS" in text, f"{name}: no mbarrier"
        assert not re.search(r"\bFFMA\b", text), f"{name}: scalar FFMA found (contraction!)"
        for l in lines:
            if "FFMA2" in l:
                # addend must be the broadcast scalar -0.0 (".F32"), never a packed travel-time pair (".F32x2")
                assert re.search(r"FFMA2 R\d+, .*, U?R\d+(\.reuse)?\.F32 ;", l), f"{name}: FFMA2 with a packed addend: {l}"
    stock = [k for k in relax if "MaskListIJLj" in k]
    assert len(stock) >= 3
    for name in stock:
        text = "\n".join(relax[name])
        assert "FADD2" in text and "FMNMX3" in text, f"{name}: packed fp32x2 path missing"


def test_simple_and_verifier_kernels_do_not_fuse_either():
    for name, lines in _sass().items():
        if "relax_simple" in name or "count_violations" in name:
            assert not re.search(r"\bFFMA2?\b", "\n".join(lines)), name
