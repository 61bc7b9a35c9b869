"""GPU (-m gpu): the CUDA path through the C ABI against the reference-generated goldens and
the CPU oracle.  Bar: BIT-EXACT float32 fields (BASELINE.json north_star)."""
import itertools
import os
import subprocess

import numpy as np
import pytest

import oracle
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import api, workloads as W

from conftest import ROOT, assert_bit_equal, make_field

pytestmark = pytest.mark.gpu

KERNELS = [api.KERNEL_SIMPLE, api.KERNEL_TILED]
LOOPS = [api.LOOP_BATCHED, api.LOOP_GRAPH]


def test_extension_is_the_thing_that_runs():
    assert P.lib_path().exists()
    assert P.device_count() >= 1


@pytest.mark.parametrize("kernel,loop", list(itertools.product(KERNELS, LOOPS)))
def test_goldens_bit_exact(golden_small, kernel, loop):
    by_case = {}
    for m in golden_small:
        by_case.setdefault(m["case"], []).append(m)
    for case, ms in by_case.items():
        v = make_field(ms[0]["kind"], ms[0]["dims"], ms[0]["seed"])
        starts = [m["start"] for m in ms]
        tt, st = P.solve(v, W.star(ms[0]["star"]), starts, kernel=kernel, loop=loop)
        assert st.kernel_used == kernel
        assert st.kernel_launches > 0 and st.relaxations > 0
        for s, m in enumerate(ms):
            assert_bit_equal(tt[s], m["tt"], f"{case}[{s}] kernel={kernel} loop={loop}")


@pytest.mark.parametrize("rxy", ["4", "7"])
def test_small_star_in_wider_halo_variants(rxy, golden_small, monkeypatch):
    """3-FS normally runs the RXY=2 instantiation; force the wider ones over the same input."""
    monkeypatch.setenv("SWEEPTT_FORCE_RXY", rxy)
    m = golden_small[0]
    v = make_field(m["kind"], m["dims"], m["seed"])
    with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
        ctx.set_model(v); ctx.set_star(W.star("3")); ctx.set_sources([m["start"]])
        ctx.run()
        assert_bit_equal(ctx.get_tt(0), m["tt"])


@pytest.mark.parametrize("axis", ["0", "1", "2"])
def test_every_window_axis_permutation(axis, golden_small, monkeypatch):
    """The solver permutes axes so the register-window axis tiles best; force each choice."""
    monkeypatch.setenv("SWEEPTT_WINDOW_AXIS", axis)
    for m in golden_small:
        if m["idx"] != 0:
            continue
        v = make_field(m["kind"], m["dims"], m["seed"])
        with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
            ctx.set_model(v); ctx.set_star(W.star(m["star"])); ctx.set_sources([m["start"]])
            ctx.run()
            assert ctx.count_violations(0) == 0
            assert_bit_equal(ctx.get_tt(0), m["tt"], f"{m['case']} window axis {axis}")


@pytest.mark.parametrize("dims", [(1, 1, 1), (1, 9, 1), (8, 8, 32), (9, 9, 33), (7, 7, 31), (16, 8, 64), (3, 40, 5)])
def test_degenerate_and_tile_edge_shapes(dims):
    v = W.random_field(dims, seed=sum(dims))
    off = W.star("818")
    corners = sorted({(0, 0, 0), (dims[0] - 1, dims[1] - 1, dims[2] - 1), (dims[0] // 2, dims[1] // 2, dims[2] // 2)})
    for kernel in KERNELS:
        tt, _ = P.solve(v, off, corners, kernel=kernel)
        for s, p in enumerate(corners):
            ref, _, _ = oracle.solve(v, off, p)
            assert_bit_equal(tt[s], ref, f"dims={dims} start={p} kernel={kernel}")


def test_quirk_nodes_in_and_out_of_bounds():
    """start - o_last inside the box, on its face, and outside it."""
    v = W.constant_field((20, 17, 13))
    off = W.star("818")  # o_last = (7,1,1)
    starts = [(12, 9, 11), (7, 1, 1), (6, 9, 11), (19, 16, 12), (7, 0, 5)]
    tt, _ = P.solve(v, off, starts, kernel=api.KERNEL_TILED)
    for s, p in enumerate(starts):
        ref, _, _ = oracle.solve(v, off, p)
        assert_bit_equal(tt[s], ref, f"start={p}")


def test_batched_sources_with_different_convergence_times():
    v = W.contrast_field((30, 26, 40), seed=9)
    off = W.star("5")
    starts = [(0, 0, 0), (15, 13, 20), (29, 25, 39), (15, 13, 21), (1, 24, 3), (29, 0, 0), (14, 14, 39)]
    tt, st = P.solve(v, off, starts, kernel=api.KERNEL_TILED)
    for s, p in enumerate(starts):
        ref, _, _ = oracle.solve(v, off, p)
        assert_bit_equal(tt[s], ref, f"source {s}")
    assert st.tile_visits > 0


def test_asymmetric_star_on_gpu():
    rng = np.random.default_rng(5)
    v = W.random_field((18, 12, 20), seed=2)
    full = W.star("5")
    off = full[np.sort(rng.permutation(len(full))[:150])]
    for kernel in KERNELS:
        tt, _ = P.solve(v, off, [(9, 6, 10), (0, 11, 19)], kernel=kernel)
        for s, p in enumerate([(9, 6, 10), (0, 11, 19)]):
            ref, _, _ = oracle.solve(v, off, p)
            assert_bit_equal(tt[s], ref, f"asymmetric kernel={kernel}")


def test_wide_star_falls_back_to_simple_kernel():
    off = np.array([[9, 0, 0], [-9, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1, 1, 1], [-1, -1, -1]], np.int32)
    v = W.random_field((25, 6, 6), seed=3)
    tt, st = P.solve(v, off, [(12, 3, 3)])
    assert st.kernel_used == api.KERNEL_SIMPLE
    ref, _, _ = oracle.solve(v, off, (12, 3, 3))
    assert_bit_equal(tt[0], ref)
    with pytest.raises(P.SweepError):
        P.solve(v, off, [(12, 3, 3)], kernel=api.KERNEL_TILED)


def test_context_step_violations_and_restart_from_upper_bounds():
    v = W.heterogeneous_field((24, 20, 36), seed=4)
    off = W.star("818")
    p = (12, 10, 35)
    ref, _, _ = oracle.solve(v, off, p)
    with P.SweepContext(kernel=api.KERNEL_TILED) as ctx:
        ctx.set_model(v); ctx.set_star(off); ctx.set_sources([p])
        ctx.reset()
        changed, _ = ctx.step(1)
        assert changed and ctx.count_violations(0) > 0
        st = ctx.run()
        assert ctx.count_violations(0) == 0
        assert_bit_equal(ctx.get_tt(0), ref)
        assert st.relaxations > 0 and ctx.relaxations_per_round > 0
        # restart from a valid upper bound (a partially swept oracle state): same fixed point
        part = oracle.init_tt(v.shape, p)
        oracle.sweep(v, part, off, p)
        ctx.reset()
        ctx.put_tt(0, part)
        while ctx.step(4)[0]:
            pass
        assert_bit_equal(ctx.get_tt(0), ref)
        assert ctx.pool_bytes > 0


def test_errors_are_loud():
    v = W.random_field((8, 8, 8))
    with pytest.raises(P.SweepError, match="outside"):
        P.solve(v, W.star("3"), [(8, 0, 0)])
    with pytest.raises(P.SweepError):
        P.solve(v, W.star("3")[:1], [(0, 0, 0)])


def test_idempotent_and_deterministic():
    v = W.heterogeneous_field((33, 25, 40), seed=5)
    off = W.star("818")
    a, _ = P.solve(v, off, [(16, 12, 39)])
    b, _ = P.solve(v, off, [(16, 12, 39)])
    assert_bit_equal(a, b)


def test_cli_drop_in(tmp_path):
    v = W.random_field((12, 11, 9), seed=1)
    off, starts = W.star("5"), [(6, 2, 8), (0, 10, 0)]
    P.vbox_store(tmp_path / "v.vbox", v, origin=(1, 1, 1))
    W.write_star_file(tmp_path / "fs.txt", off)
    W.write_start_file(tmp_path / "start.txt", starts)
    exe = P.lib_path().parent / "sweep-tt-multistart"
    env = dict(os.environ, SWEEPTT_TT_BIN=str(tmp_path / "tt.bin"))
    r = subprocess.run([str(exe), "v.vbox", "fs.txt", "start.txt"], cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.splitlines()
    assert out[0] == "Loading velocity model file: v.vbox... done."
    assert out[1] == "Velocity model dimensions: 12 x 11 x 9"
    assert "Forward star size: 422" in out and "starting point 1: 0 10 0" in out
    assert "numradius: 4, fsindex[3]: 218" in out
    ref = np.stack([oracle.solve(v, off, p)[0] for p in starts])
    got = np.fromfile(tmp_path / "tt.bin", np.float32).reshape(ref.shape)
    assert_bit_equal(got, ref)
    oracle.write_output_tt(tmp_path / "expected.tt", ref)
    assert (tmp_path / "output.tt").read_bytes() == (tmp_path / "expected.tt").read_bytes()
    # unopenable inputs: message + exit(1), like serial_new/...c:81-84
    r = subprocess.run([str(exe), "missing.vbox", "fs.txt", "start.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open velocity model file: missing.vbox" in r.stdout
