"""Input fixtures and synthetic workloads (BASELINE.json configs).

The reference's two 241x241x51 velocity files are missing blobs (.MISSING_LARGE_BLOBS), so
every config runs on a synthetic box with a recorded seed, written through the same file
formats.  Forward stars and start points are the reference's own input data
(docs/{3,5,818}-FS.txt, docs/start-*-241-241-51.txt), packed in data/inputs.npz by
tools/pack_inputs.py.
"""
from __future__ import annotations

import functools
import pathlib

import numpy as np

_ROOT = pathlib.Path(__file__).resolve().parents[1]


@functools.lru_cache(maxsize=None)
def _inputs():
    with np.load(_ROOT / "data" / "inputs.npz") as z:
        return {k: z[k].copy() for k in z.files}


def star(name: str) -> np.ndarray:
    """'3', '5' or '818' -> int32[L,3] offsets in file order (order matters: the last entry is unused)."""
    return _inputs()[f"fs_{name}"]


def starts(n: int) -> np.ndarray:
    """1, 4, 10, 24 or 111 -> int32[n,3] zero-based start points for the 241x241x51 box."""
    return _inputs()[f"start_{n}"]


def write_star_file(path, offsets):
    off = np.asarray(offsets).reshape(-1, 3)
    with open(path, "w") as f:
        f.write(f"{len(off)}\n")
        for a, b, c in off:
            f.write(f"{a} {b} {c}\n")


def write_start_file(path, pts):
    pts = np.asarray(pts).reshape(-1, 3)
    with open(path, "w") as f:
        f.write(f"{len(pts)}\n")
        for a, b, c in pts:
            f.write(f"{a} {b} {c}\n")


def write_text_a(path, v, origin=(1, 1, 1)):
    """Text dialect A: 'x,y,z,v' per line, 1-based (include/velocityboxfiler.h:95-103)."""
    nx, ny, nz = v.shape
    with open(path, "w") as f:
        for x in range(nx):
            for y in range(ny):
                for z in range(nz):
                    f.write(f"{x + origin[0]},{y + origin[1]},{z + origin[2]},{float(v[x, y, z])!r}\n")


def write_text_b(path, v):
    """Text dialect B: 'nx ny nz' then bare floats (old/wavefront-openmp/wave-multistart.c:151-161)."""
    with open(path, "w") as f:
        f.write("%d %d %d\n" % v.shape)
        for val in v.ravel():
            f.write(f"{float(val)!r}\n")


def constant_field(dims=(241, 241, 51), value=0.25) -> np.ndarray:
    """Config 1 stand-in: constant slowness."""
    return np.full(dims, value, np.float32)


def heterogeneous_field(dims=(241, 241, 51), seed=7) -> np.ndarray:
    """Configs 2-5 stand-in for velocity-241-241-51-nonConst: depth gradient 0.30 -> 0.15 along z
    times (1 +- 10 %) uniform noise, numpy default_rng(seed), float32 throughout."""
    nx, ny, nz = dims
    rng = np.random.default_rng(seed)
    grad = np.linspace(0.30, 0.15, nz, dtype=np.float32)
    noise = rng.random(dims, dtype=np.float32) * np.float32(0.2) + np.float32(0.9)
    return (grad[None, None, :] * noise).astype(np.float32)


def random_field(dims, seed=0, lo=0.1, hi=0.4) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.random(dims, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def contrast_field(dims, seed=0) -> np.ndarray:
    """High-contrast blocks (x20 slowness jumps): stresses late corrections behind the front."""
    rng = np.random.default_rng(seed)
    coarse = rng.choice(np.array([0.02, 0.1, 0.4], np.float32), size=tuple((d + 3) // 4 for d in dims))
    v = np.kron(coarse, np.ones((4, 4, 4), np.float32))[: dims[0], : dims[1], : dims[2]]
    return np.ascontiguousarray(v, dtype=np.float32)


def visits_per_sweep(dims, offsets, used=None) -> int:
    """In-bounds (node, offset) visits of one reference sweep (serial_new/...c:203-214)."""
    off = np.asarray(offsets).reshape(-1, 3)
    if used is None:
        used = len(off) - 1
    n = 0
    for a, b, c in off[:used]:
        ex, ey, ez = dims[0] - abs(int(a)), dims[1] - abs(int(b)), dims[2] - abs(int(c))
        if ex > 0 and ey > 0 and ez > 0:
            n += ex * ey * ez
    return n
