"""CPU: the host-side edge-set analysis (product code, csrc/pullstar.cpp through the C ABI).
A numpy float32 Jacobi over the pull star it returns must reach the reference's field."""
import numpy as np
import pytest

import oracle
import uoparallel_seismic_project_b200 as P
from uoparallel_seismic_project_b200 import workloads as W

from conftest import assert_bit_equal, make_field


def pull_jacobi(v, ijk, hd, guard, start, max_iter=500):
    """tt[n] = min(tt[n], fl(fl(hd*fl(v_n+v_m)) + tt[m])) for every pull; float32 throughout."""
    nx, ny, nz = v.shape
    tt = np.full(v.shape, np.inf, np.float32)
    tt[tuple(start)] = 0
    for it in range(max_iter):
        new = tt.copy()
        for (a, b, c), h, g in zip(ijk, hd, guard):
            xs = slice(max(0, -a), min(nx, nx - a)); xm = slice(max(0, a), min(nx, nx + a))
            ys = slice(max(0, -b), min(ny, ny - b)); ym = slice(max(0, b), min(ny, ny + b))
            zs = slice(max(0, -c), min(nz, nz - c)); zm = slice(max(0, c), min(nz, nz + c))
            if xs.start >= xs.stop or ys.start >= ys.stop or zs.start >= zs.stop:
                continue
            cand = (np.float32(h) * (v[xs, ys, zs] + v[xm, ym, zm])).astype(np.float32) + tt[xm, ym, zm]
            if g:  # invalid where the neighbour is the start point
                n = np.array(start) - np.array([a, b, c])
                if xs.start <= n[0] < xs.stop and ys.start <= n[1] < ys.stop and zs.start <= n[2] < zs.stop:
                    cand[n[0] - xs.start, n[1] - ys.start, n[2] - zs.start] = np.inf
            np.minimum(new[xs, ys, zs], cand, out=new[xs, ys, zs])
        new[tuple(start)] = 0
        if np.array_equal(new, tt):
            return tt, it + 1
        tt = new
    raise AssertionError("no convergence")


def test_symmetric_stars_have_one_guarded_pull():
    for name, last in (("3", (2, 2, 1)), ("5", (4, 3, 0)), ("818", (7, 1, 1))):
        off = W.star(name)
        ijk, hd, guard = P.build_pull_star(off)
        assert len(ijk) == len(off)
        assert guard.sum() == 1 and tuple(ijk[guard == 1][0]) == last == tuple(off[-1])
        d = oracle.star_distances(off)
        lut = {tuple(o): x for o, x in zip(off, d)}
        for o, h in zip(ijk, hd):
            assert np.float32(lut[tuple(o)] * np.float32(0.5)) == h


def test_pull_form_reaches_reference_field(golden_small):
    for m in golden_small:
        if int(np.prod(m["dims"])) > 6000:
            continue
        v = make_field(m["kind"], m["dims"], m["seed"])
        ijk, hd, guard = P.build_pull_star(W.star(m["star"]))
        tt, _ = pull_jacobi(v, ijk, hd, guard, m["start"])
        assert_bit_equal(tt, m["tt"], f"{m['case']}[{m['idx']}]")


def test_asymmetric_and_truncated_stars():
    """Stars that are not symmetric (or whose negations are missing) exercise the guarded set."""
    rng = np.random.default_rng(0)
    v = W.random_field((9, 8, 7), seed=11)
    full = W.star("3")
    for trial in range(4):
        sel = rng.permutation(len(full))[: 30 + 10 * trial]
        off = full[np.sort(sel)]
        start = (4, 4, 3)
        ref, _, _ = oracle.solve(v, off, start)
        ijk, hd, guard = P.build_pull_star(off)
        tt, _ = pull_jacobi(v, ijk, hd, guard, start)
        assert_bit_equal(tt, ref, f"asymmetric trial {trial}")


def test_star_used_override_and_zero_offset():
    off = np.vstack([W.star("3")[:20], [[0, 0, 0]], W.star("3")[20:40]])
    ijk, hd, guard = P.build_pull_star(off)
    assert not any((o == 0).all() for o in ijk)
    ijk_all, _, guard_all = P.build_pull_star(W.star("3"), star_used=98)
    assert guard_all.sum() == 0 and len(ijk_all) == 98


@pytest.mark.parametrize("name,nw", [("818", 16), ("5", 16), ("3", 4), ("3", 16)])
def test_column_split_covers_every_column_once_and_balances_the_warps(name, nw):
    """The tiled kernel shares a star's columns out between its warps with cut points computed on the host
    (csrc/pullstar.cpp split_columns): every table must partition every pattern group, and the warps' total
    costs (popcount + 1.5 per column, plus the head starts of the owner / feeder / finisher warps) must be level."""
    import uoparallel_seismic_project_b200 as P
    from uoparallel_seismic_project_b200 import api, workloads as W
    kmasks, cuts = api.column_split(W.star(name), nw)
    cost = np.array([bin(int(m)).count("1") + 1.5 for m in kmasks])
    gbeg = [0] + [i for i in range(1, len(kmasks)) if kmasks[i] != kmasks[i - 1]] + [len(kmasks)]
    assert cuts.shape == (6, len(gbeg) - 1, nw + 1)
    bias = [(8.0, 2.0, 14.0), (10.0, 2.0, 60.0)]  # csrc/pullstar.h kDefaultBias
    for t in range(6):
        parts = nw if t % 3 == 0 else nw // 2
        warp0 = nw // 2 if t % 3 == 2 else 0
        load = np.zeros(parts)
        for pt in range(parts):
            w = warp0 + pt
            load[pt] += bias[t // 3][0] * (pt == 0) + bias[t // 3][1] * (w == nw // 2 - 1) + bias[t // 3][2] * (w == nw - 1)
        for g in range(len(gbeg) - 1):
            row = cuts[t, g].astype(int)
            assert row[0] == gbeg[g] and row[parts] == gbeg[g + 1], (t, g)
            assert all(row[p] <= row[p + 1] for p in range(nw)), (t, g)
            assert all(row[p] == gbeg[g + 1] for p in range(parts, nw + 1)), (t, g)
            for pt in range(parts):
                load[pt] += cost[row[pt]:row[pt + 1]].sum()
        # level: no warp more than one heavy column (16.5 cost units) above the mean -- where there is enough
        # work to level at all (a small star forced into the 16-warp kernel has less cost than head starts)
        if cost.sum() / parts > 60:
            assert load.max() - load.mean() <= 16.5 + 1e-9, (name, nw, t, load)
