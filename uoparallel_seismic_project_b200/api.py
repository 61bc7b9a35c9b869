"""ctypes binding of include/sweeptt.h -- the reference-facing call surface.

Names follow the reference: ``FS`` / ``START`` are the structs of
serial_new/sweep-tt-multistart.c:46-58; :func:`solve` stands where ``cudaRun`` /
the ``while(anychange) sweepXYZ`` loop stood (cuda/cudasweep-tt-multistart.cu:227,
serial_new/...c:150-170); boxes are FLOATBOX-ordered ``float32[nx, ny, nz]`` (z fastest,
include/floatbox.h:127-129).  Error behaviour mirrors the ABI: a zero return becomes a
:class:`SweepError` carrying ``sweeptt_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
from dataclasses import dataclass

import numpy as np

_PKG = pathlib.Path(__file__).resolve().parent
_LIB = None

KERNEL_AUTO, KERNEL_SIMPLE, KERNEL_TILED = 0, 1, 2
LOOP_AUTO, LOOP_BATCHED, LOOP_GRAPH = 0, 1, 2


class SweepError(RuntimeError):
    pass


class FS(C.Structure):
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("k", C.c_int), ("d", C.c_float)]


class START(C.Structure):
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("k", C.c_int)]


class _Opts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int), ("device", C.c_int), ("num_devices", C.c_int), ("kernel", C.c_int),
        ("loop", C.c_int), ("rounds_per_poll", C.c_int), ("max_rounds", C.c_int), ("star_used", C.c_int),
        ("verbose", C.c_int), ("slab_axis", C.c_int), ("profile_kernels", C.c_int),
    ]


class _Stats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int), ("rounds", C.c_int), ("kernel_used", C.c_int), ("devices_used", C.c_int),
        ("kernel_launches", C.c_longlong), ("tile_visits", C.c_longlong), ("relaxations", C.c_longlong),
        ("solve_ms", C.c_double), ("relax_kernel_ms", C.c_double), ("relax_launches", C.c_longlong),
        ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
        ("units_run", C.c_longlong), ("units_changed", C.c_longlong),
    ]


@dataclass
class SweepStats:
    rounds: int = 0
    kernel_used: int = 0
    devices_used: int = 0
    kernel_launches: int = 0
    tile_visits: int = 0
    relaxations: int = 0
    solve_ms: float = 0.0
    relax_kernel_ms: float = 0.0
    relax_launches: int = 0
    h2d_ms: float = 0.0
    d2h_ms: float = 0.0
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    units_run: int = 0
    units_changed: int = 0

    @classmethod
    def _from(cls, s: _Stats) -> "SweepStats":
        return cls(**{f: getattr(s, f) for f in cls.__dataclass_fields__})


def lib_path() -> pathlib.Path:
    import os
    if os.environ.get("SWEEPTT_LIB"):  # developer override: an experimental build of the same ABI
        return pathlib.Path(os.environ["SWEEPTT_LIB"])
    return _PKG / "lib" / "libsweeptt.so"


_EXPORTS = [
    "sweeptt_device_count", "sweeptt_device_info", "sweeptt_last_error", "sweeptt_version",
    "sweeptt_star_fill_distances", "sweeptt_build_pull_star", "sweeptt_debug_column_split", "sweeptt_solve", "sweeptt_release_cache",
    "sweeptt_host_alloc", "sweeptt_host_free",
    "sweeptt_create", "sweeptt_destroy", "sweeptt_set_stream", "sweeptt_set_model", "sweeptt_set_star",
    "sweeptt_set_sources", "sweeptt_run", "sweeptt_step", "sweeptt_reset", "sweeptt_get_tt", "sweeptt_put_tt",
    "sweeptt_count_violations", "sweeptt_relaxations_per_round", "sweeptt_pool_bytes", "sweeptt_tiles_per_source", "sweeptt_solve_slabs",
    "sweeptt_solve_slabs_vbox", "sweeptt_vbox_dims",
    "sweeptt_vbox_load", "sweeptt_vbox_store", "sweeptt_vbox_load_subset", "sweeptt_text_load",
    "sweeptt_star_load", "sweeptt_starts_load", "sweeptt_write_output_tt", "sweeptt_free",
]


def load_library() -> C.CDLL:
    """Load libsweeptt.so (never a fallback: a missing extension is an error)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise SweepError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(the sweep has no CPU fallback)")
    lib = C.CDLL(str(p))
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.sweeptt_last_error.restype = C.c_char_p
    lib.sweeptt_version.restype = C.c_char_p
    lib.sweeptt_device_info.argtypes = [C.c_int, C.c_char_p, C.c_int, ip, ip, C.POINTER(C.c_size_t)]
    lib.sweeptt_star_fill_distances.argtypes = [C.POINTER(FS), C.c_int, C.c_float]
    lib.sweeptt_star_fill_distances.restype = None
    lib.sweeptt_build_pull_star.argtypes = [C.POINTER(FS), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.sweeptt_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(FS), C.c_int, C.POINTER(START),
                                  C.c_int, C.POINTER(C.c_void_p), C.POINTER(_Opts), C.POINTER(_Stats)]
    lib.sweeptt_release_cache.restype = None
    lib.sweeptt_host_alloc.argtypes = [C.c_size_t]
    lib.sweeptt_host_alloc.restype = C.c_void_p
    lib.sweeptt_host_free.argtypes = [C.c_void_p]
    lib.sweeptt_host_free.restype = None
    lib.sweeptt_create.argtypes = [C.POINTER(_Opts)]
    lib.sweeptt_create.restype = C.c_void_p
    lib.sweeptt_destroy.argtypes = [C.c_void_p]
    lib.sweeptt_destroy.restype = None
    lib.sweeptt_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.sweeptt_set_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.sweeptt_set_star.argtypes = [C.c_void_p, C.POINTER(FS), C.c_int]
    lib.sweeptt_set_sources.argtypes = [C.c_void_p, C.POINTER(START), C.c_int]
    lib.sweeptt_run.argtypes = [C.c_void_p, C.POINTER(_Stats)]
    lib.sweeptt_step.argtypes = [C.c_void_p, C.c_int, ip, C.POINTER(_Stats)]
    lib.sweeptt_reset.argtypes = [C.c_void_p]
    lib.sweeptt_get_tt.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.sweeptt_put_tt.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.sweeptt_count_violations.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
    lib.sweeptt_relaxations_per_round.argtypes = [C.c_void_p]
    lib.sweeptt_relaxations_per_round.restype = C.c_longlong
    lib.sweeptt_pool_bytes.argtypes = [C.c_void_p]
    lib.sweeptt_pool_bytes.restype = C.c_size_t
    lib.sweeptt_tiles_per_source.argtypes = [C.c_void_p]
    lib.sweeptt_tiles_per_source.restype = C.c_longlong
    lib.sweeptt_solve_slabs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(FS), C.c_int, START,
                                        C.c_void_p, C.POINTER(_Opts), C.POINTER(_Stats)]
    lib.sweeptt_solve_slabs_vbox.argtypes = [C.c_char_p, C.POINTER(FS), C.c_int, START, C.c_void_p, C.POINTER(_Opts),
                                             C.POINTER(_Stats)]
    lib.sweeptt_vbox_dims.argtypes = [C.c_char_p, ip]
    lib.sweeptt_vbox_load.argtypes = [C.c_char_p, C.POINTER(fp), ip, ip]
    lib.sweeptt_vbox_store.argtypes = [C.c_char_p, C.c_void_p, ip, ip]
    lib.sweeptt_vbox_load_subset.argtypes = [C.c_char_p, ip, ip, C.POINTER(fp)]
    lib.sweeptt_text_load.argtypes = [C.c_char_p, C.POINTER(fp), ip, ip]
    lib.sweeptt_star_load.argtypes = [C.c_char_p, C.c_float, C.POINTER(C.POINTER(FS)), ip]
    lib.sweeptt_starts_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(START)), ip]
    lib.sweeptt_write_output_tt.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
    lib.sweeptt_free.argtypes = [C.c_void_p]
    lib.sweeptt_free.restype = None
    _LIB = lib
    return lib


def _check(ok, what: str):
    if not ok:
        raise SweepError(f"{what}: {load_library().sweeptt_last_error().decode()}")


def device_count() -> int:
    return load_library().sweeptt_device_count()


def make_star(offsets, delta: float = 10.0):
    """(L,3) int offsets -> FS[L] with d filled the reference's way (serial_new/...c:120-128)."""
    off = np.asarray(offsets, dtype=np.int32).reshape(-1, 3)
    arr = (FS * len(off))()
    for l, (a, b, c) in enumerate(off):
        arr[l].i, arr[l].j, arr[l].k = int(a), int(b), int(c)
    load_library().sweeptt_star_fill_distances(arr, len(off), C.c_float(delta))
    return arr


def _make_starts(starts):
    st = np.asarray(starts, dtype=np.int32).reshape(-1, 3)
    arr = (START * len(st))()
    for s, (a, b, c) in enumerate(st):
        arr[s].i, arr[s].j, arr[s].k = int(a), int(b), int(c)
    return arr


def _as_star(star, delta):
    return star if isinstance(star, C.Array) and star._type_ is FS else make_star(star, delta)


def _opts(**kw) -> _Opts:
    o = _Opts()
    o.struct_size = C.sizeof(_Opts)
    o.device = -1
    for k, v in kw.items():
        if v is not None:
            setattr(o, k, v)
    return o


def build_pull_star(star, star_used: int = 0, delta: float = 10.0):
    """Host-side edge-set analysis: (offsets int32[P,3], half distances f32[P], guard int32[P])."""
    fs = _as_star(star, delta)
    lib = load_library()
    cap = 2 * len(fs) + 4
    ijk = np.zeros((cap, 3), np.int32)
    hd = np.zeros(cap, np.float32)
    gd = np.zeros(cap, np.int32)
    n = lib.sweeptt_build_pull_star(fs, len(fs), star_used, ijk.ctypes.data, hd.ctypes.data, gd.ctypes.data, cap)
    if n < 0:
        raise SweepError("sweeptt_build_pull_star failed")
    return ijk[:n].copy(), hd[:n].copy(), gd[:n].copy()


def column_split(star, nw: int, delta: float = 10.0):
    """Test hook: (kmasks uint32[C], cuts uint16[6, G, nw+1]) -- how the tiled kernel shares the star's columns
    out between `nw` warps (six tables, see include/sweeptt.h)."""
    fs = _as_star(star, delta)
    lib = load_library()
    lib.sweeptt_debug_column_split.argtypes = [C.POINTER(FS), C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p,
                                               C.c_int, C.c_void_p, C.c_int]
    lib.sweeptt_debug_column_split.restype = C.c_int
    cuts = np.zeros(6 * 64 * (nw + 1), np.uint16)
    kmasks = np.zeros(4096, np.uint32)
    ng = C.c_int(0)
    n = lib.sweeptt_debug_column_split(fs, len(fs), nw, C.byref(ng), cuts.ctypes.data, cuts.size, kmasks.ctypes.data,
                                       kmasks.size)
    if n < 0:
        raise SweepError("sweeptt_debug_column_split failed")
    return kmasks[:n].copy(), cuts[: 6 * ng.value * (nw + 1)].reshape(6, ng.value, nw + 1).copy()


def solve(slowness, star, starts, *, delta: float = 10.0, out=None, device: int | None = None,
          num_devices: int | None = None, kernel: int | None = None, loop: int | None = None,
          max_rounds: int | None = None, star_used: int | None = None, profile_kernels: int | None = None,
          rounds_per_poll: int | None = None, verbose: int | None = None):
    """One-shot multi-start solve with HOST buffers (the `cudaRun` replacement).

    slowness: float32[nx,ny,nz]; star: (L,3) offsets or FS array; starts: (S,3).
    Returns (tt float32[S,nx,ny,nz], SweepStats).  `out` may be a preallocated (pinned) array.
    """
    lib = load_library()
    v = np.ascontiguousarray(slowness, dtype=np.float32)
    if v.ndim != 3:
        raise SweepError("slowness must be a 3-D box")
    nx, ny, nz = v.shape
    fs = _as_star(star, delta)
    st = _make_starts(starts)
    ns = len(st)
    if out is None:
        out = np.empty((ns, nx, ny, nz), np.float32)
    if out.shape != (ns, nx, ny, nz) or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise SweepError("out must be C-contiguous float32[S,nx,ny,nz]")
    ptrs = (C.c_void_p * ns)(*[out[s].ctypes.data for s in range(ns)])
    o = _opts(device=device, num_devices=num_devices, kernel=kernel, loop=loop, max_rounds=max_rounds,
              star_used=star_used, profile_kernels=profile_kernels, rounds_per_poll=rounds_per_poll, verbose=verbose)
    s = _Stats()
    _check(lib.sweeptt_solve(v.ctypes.data, nx, ny, nz, fs, len(fs), st, ns, ptrs, C.byref(o), C.byref(s)),
           "sweeptt_solve")
    return out, SweepStats._from(s)


def solve_slabs(slowness, star, start, *, num_slabs: int, slab_axis: int = 0, delta: float = 10.0,
                max_rounds: int | None = None, rounds_per_poll: int | None = None, verbose: int | None = None):
    """ONE source on ONE grid spread over `num_slabs` parts (the visible GPUs; more parts than GPUs share devices
    round-robin): the travel-time box is one range of peer memory whose pages are dealt block-cyclically along
    `slab_axis`, every part relaxes its own blocks and reads halo planes from the owner over NVLink inside the
    relaxation kernel -- no exchange step.  Replaces the MPI ghost-cell programs (mpi/16partsmpi.c:740-909).
    Returns (tt float32[nx,ny,nz], SweepStats)."""
    lib = load_library()
    v = np.ascontiguousarray(slowness, dtype=np.float32)
    nx, ny, nz = v.shape
    fs = _as_star(star, delta)
    st = START(int(start[0]), int(start[1]), int(start[2]))
    out = np.empty(v.shape, np.float32)
    o = _opts(num_devices=num_slabs, slab_axis=slab_axis, max_rounds=max_rounds, rounds_per_poll=rounds_per_poll,
              verbose=verbose)
    s = _Stats()
    _check(lib.sweeptt_solve_slabs(v.ctypes.data, nx, ny, nz, fs, len(fs), st, out.ctypes.data, C.byref(o), C.byref(s)),
           "sweeptt_solve_slabs")
    return out, SweepStats._from(s)


def solve_slabs_vbox(path, star, start, *, num_slabs: int, slab_axis: int = 0, delta: float = 10.0):
    """Like solve_slabs, but every part reads only the planes of its blocks from the .vbox file (subset loader)."""
    lib = load_library()
    d = (C.c_int * 3)()
    _check(lib.sweeptt_vbox_dims(os.fsencode(path), d), "sweeptt_vbox_dims")
    fs = _as_star(star, delta)
    st = START(int(start[0]), int(start[1]), int(start[2]))
    out = np.empty(tuple(d), np.float32)
    o = _opts(num_devices=num_slabs, slab_axis=slab_axis)
    s = _Stats()
    _check(lib.sweeptt_solve_slabs_vbox(os.fsencode(path), fs, len(fs), st, out.ctypes.data, C.byref(o), C.byref(s)),
           "sweeptt_solve_slabs_vbox")
    return out, SweepStats._from(s)


def solve_raw(v_ptr: int, dims, fs, starts_arr, out_ptrs, opts: _Opts):
    """Pointer-level call for bench.py's e2e leg (pinned torch tensors; no numpy in the way)."""
    lib = load_library()
    s = _Stats()
    _check(lib.sweeptt_solve(C.c_void_p(v_ptr), dims[0], dims[1], dims[2], fs, len(fs), starts_arr, len(starts_arr),
                             out_ptrs, C.byref(opts), C.byref(s)), "sweeptt_solve")
    return SweepStats._from(s)


class SweepContext:
    """Device-resident context: padded slowness box + pool of travel-time boxes on one GPU."""

    def __init__(self, *, device: int | None = None, kernel: int | None = None, loop: int | None = None,
                 max_rounds: int | None = None, star_used: int | None = None, profile_kernels: int | None = None,
                 rounds_per_poll: int | None = None, verbose: int | None = None):
        self._lib = load_library()
        o = _opts(device=device, kernel=kernel, loop=loop, max_rounds=max_rounds, star_used=star_used,
                  profile_kernels=profile_kernels, rounds_per_poll=rounds_per_poll, verbose=verbose)
        self._h = self._lib.sweeptt_create(C.byref(o))
        if not self._h:
            raise SweepError(f"sweeptt_create: {self._lib.sweeptt_last_error().decode()}")
        self.dims = None
        self.nsrc = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sweeptt_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int):
        _check(self._lib.sweeptt_set_stream(self._h, C.c_void_p(cuda_stream)), "sweeptt_set_stream")

    def set_model(self, slowness):
        v = np.ascontiguousarray(slowness, dtype=np.float32)
        self.dims = v.shape
        _check(self._lib.sweeptt_set_model(self._h, v.ctypes.data, *v.shape), "sweeptt_set_model")

    def set_star(self, star, delta: float = 10.0):
        fs = _as_star(star, delta)
        _check(self._lib.sweeptt_set_star(self._h, fs, len(fs)), "sweeptt_set_star")

    def set_sources(self, starts):
        st = _make_starts(starts)
        self.nsrc = len(st)
        _check(self._lib.sweeptt_set_sources(self._h, st, len(st)), "sweeptt_set_sources")

    def run(self) -> SweepStats:
        s = _Stats()
        _check(self._lib.sweeptt_run(self._h, C.byref(s)), "sweeptt_run")
        return SweepStats._from(s)

    def reset(self):
        _check(self._lib.sweeptt_reset(self._h), "sweeptt_reset")

    def step(self, rounds: int = 1):
        s = _Stats()
        ch = C.c_int(0)
        _check(self._lib.sweeptt_step(self._h, rounds, C.byref(ch), C.byref(s)), "sweeptt_step")
        return bool(ch.value), SweepStats._from(s)

    def get_tt(self, source: int, out=None):
        if out is None:
            out = np.empty(self.dims, np.float32)
        _check(self._lib.sweeptt_get_tt(self._h, source, out.ctypes.data), "sweeptt_get_tt")
        return out

    def put_tt(self, source: int, tt):
        t = np.ascontiguousarray(tt, dtype=np.float32)
        _check(self._lib.sweeptt_put_tt(self._h, source, t.ctypes.data), "sweeptt_put_tt")

    def count_violations(self, source: int) -> int:
        n = C.c_longlong(0)
        _check(self._lib.sweeptt_count_violations(self._h, source, C.byref(n)), "sweeptt_count_violations")
        return n.value

    @property
    def relaxations_per_round(self) -> int:
        return self._lib.sweeptt_relaxations_per_round(self._h)

    @property
    def tiles_per_source(self) -> int:
        return self._lib.sweeptt_tiles_per_source(self._h)

    @property
    def pool_bytes(self) -> int:
        return self._lib.sweeptt_pool_bytes(self._h)


# ---- file formats -------------------------------------------------------------------------

def _take_floats(ptr, dims):
    n = int(dims[0]) * int(dims[1]) * int(dims[2])
    arr = np.ctypeslib.as_array(ptr, shape=(n,)).reshape(tuple(int(d) for d in dims)).copy()
    load_library().sweeptt_free(ptr)
    return arr


def vbox_load(path):
    lib = load_library()
    p = C.POINTER(C.c_float)()
    o, d = (C.c_int * 3)(), (C.c_int * 3)()
    _check(lib.sweeptt_vbox_load(os.fsencode(path), C.byref(p), o, d), "sweeptt_vbox_load")
    return _take_floats(p, d), tuple(o), tuple(d)


def vbox_store(path, slowness, origin=(0, 0, 0)):
    v = np.ascontiguousarray(slowness, dtype=np.float32)
    o, d = (C.c_int * 3)(*origin), (C.c_int * 3)(*v.shape)
    _check(load_library().sweeptt_vbox_store(os.fsencode(path), v.ctypes.data, o, d), "sweeptt_vbox_store")


def vbox_load_subset(path, sub_origin, sub_dims):
    p = C.POINTER(C.c_float)()
    o, d = (C.c_int * 3)(*sub_origin), (C.c_int * 3)(*sub_dims)
    _check(load_library().sweeptt_vbox_load_subset(os.fsencode(path), o, d, C.byref(p)), "sweeptt_vbox_load_subset")
    return _take_floats(p, d)


def text_load(path):
    p = C.POINTER(C.c_float)()
    o, d = (C.c_int * 3)(), (C.c_int * 3)()
    _check(load_library().sweeptt_text_load(os.fsencode(path), C.byref(p), o, d), "sweeptt_text_load")
    return _take_floats(p, d), tuple(o), tuple(d)


def star_load(path, delta: float = 10.0):
    lib = load_library()
    p = C.POINTER(FS)()
    n = C.c_int(0)
    _check(lib.sweeptt_star_load(os.fsencode(path), C.c_float(delta), C.byref(p), C.byref(n)), "sweeptt_star_load")
    off = np.array([(p[l].i, p[l].j, p[l].k) for l in range(n.value)], np.int32)
    d = np.array([p[l].d for l in range(n.value)], np.float32)
    lib.sweeptt_free(p)
    return off, d


def starts_load(path):
    lib = load_library()
    p = C.POINTER(START)()
    n = C.c_int(0)
    _check(lib.sweeptt_starts_load(os.fsencode(path), C.byref(p), C.byref(n)), "sweeptt_starts_load")
    st = np.array([(p[s].i, p[s].j, p[s].k) for s in range(n.value)], np.int32)
    lib.sweeptt_free(p)
    return st


def write_output_tt(path, tt):
    t = np.ascontiguousarray(tt, dtype=np.float32)
    ns, nx, ny, nz = t.shape
    ptrs = (C.c_void_p * ns)(*[t[s].ctypes.data for s in range(ns)])
    _check(load_library().sweeptt_write_output_tt(os.fsencode(path), ptrs, ns, nx, ny, nz), "sweeptt_write_output_tt")
