#!/usr/bin/env python3
"""bench.py -- the driver's benchmark contract for the multi-start forward-star sweep.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 241x241x51 heterogeneous slowness box (synthetic stand-in,
seed 7: the reference's velocity files are missing blobs), docs/818-FS.txt,
docs/start-4-241-241-51.txt.  One *step* = one complete multi-start solve of that batch
(travel times re-initialised, relaxed to the fixed point).  N>1 is weak scaling: every rank
solves 4 sources of its own (rank 0: start-4; rank r: the next rows of start-111) on its own
replica of the box -- no data-path collective.

Prints ONE JSON line (rank 0).  metric = GRelax/s: in-bounds (node, offset, source) pull
evaluations executed per second, summed over GPUs; converged sources/s rides along.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time
import pathlib

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DIMS = (241, 241, 51)
WORKLOAD = "241x241x51 heterogeneous slowness (synthetic, seed 7), 818-FS, start-4 (4 sources per GPU)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4"],
                    help="config2 (default, the contract's workload): 241 box, 4 sources per GPU, weak scaling; "
                         "config3: 241 box, start-111 sharded over the GPUs (strong); "
                         "config4: 1201x1201x251, 24 sources sharded over the GPUs (strong, device-resident only)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region: NVML polled every 2 ms from a thread
    (the timed region of config 2 is ~0.3 s, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.thread = None
        self.rows = []          # (sm MHz, max MHz, watts, reasons bitmask)
        self._stop = False
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = int(get_reasons(self.h)) if get_reasons else 0
                self.rows.append((sm, self.mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def _stop_nvml(self):
        self._stop = True
        self.thread.join(timeout=2)
        nv = self.nv
        def bit(name, default):
            return getattr(nv, "nvmlClocksEventReason" + name, getattr(nv, "nvmlClocksThrottleReason" + name, default))
        bits = {"hw_slowdown": bit("HwSlowdown", 0x8), "hw_thermal_slowdown": bit("HwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": bit("SwThermalSlowdown", 0x20), "sw_power_cap": bit("SwPowerCap", 0x4)}
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [r[0] for r in self.rows]; pw = [r[2] for r in self.rows]
        reasons = sorted(n for n, b in bits.items() if any(r[3] & b for r in self.rows))
        load = [c for c, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw), "source": "nvml, 2 ms period"}

    def stop(self):
        if self.thread is not None:
            return self._stop_nvml()
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.p.kill()
            out = self.p.communicate()[0]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [c for c, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw), "source": "nvidia-smi -lms 50"}


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own serial code on the host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process per source (the mpi/backup.c scheme): warm the state with `pre` slices, then
    time `slices` calls of the reference's sweepXYZ over consecutive slices of the star."""
    src, slices, kind = args
    import numpy as np
    import oracle
    from uoparallel_seismic_project_b200 import workloads as W
    v = W.heterogeneous_field(DIMS, 7)
    off = np.ascontiguousarray(W.star("818"), np.int32)
    L = len(off) - 1
    nslice = 8
    bounds = [L * i // nslice for i in range(nslice + 1)]
    out = []
    if kind == "reference":
        lib = oracle.reference()
        lib.refh_set_star(ctypes.c_void_p(off.ctypes.data), len(off))
        lib.refh_set_velocity(ctypes.c_void_p(v.ctypes.data), *DIMS)
        lib.refh_init_source(0, int(src[0]), int(src[1]), int(src[2]))
        for k in slices:
            a, b = bounds[k % nslice], bounds[k % nslice + 1]
            t0 = time.perf_counter()
            lib.sweepXYZ(DIMS[0], DIMS[1], DIMS[2], 0, a, b)   # the reference's own function
            out.append((time.perf_counter() - t0, oracle.visits_per_sweep(DIMS, off[a:b], used=b - a)))
    else:
        lib = oracle.restatement()
        d = oracle.star_distances(off)
        tt = oracle.init_tt(DIMS, src)
        for k in slices:
            a, b = bounds[k % nslice], bounds[k % nslice + 1]
            t0 = time.perf_counter()
            lib.oracle_sweep(ctypes.c_void_p(v.ctypes.data), ctypes.c_void_p(tt.ctypes.data), DIMS[0], DIMS[1],
                             DIMS[2], ctypes.c_void_p(off[a:b].ctypes.data), ctypes.c_void_p(d[a:b].ctypes.data),
                             b - a, int(src[0]), int(src[1]), int(src[2]))
            out.append((time.perf_counter() - t0, oracle.visits_per_sweep(DIMS, off[a:b], used=b - a)))
    return out


def cpu_reference_run(steps, warmup):
    """Returns (per-step seconds (max over the source processes), visits per step, kind, cores)."""
    import multiprocessing as mp
    import oracle
    from uoparallel_seismic_project_b200 import workloads as W
    kind = "reference" if oracle.reference() is not None else "port"
    if kind == "port":
        oracle.restatement()
    starts = W.starts(4)
    cores = min(len(starts), os.cpu_count() or 1)
    slices = list(range(warmup + steps))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(tuple(int(c) for c in s), slices, kind) for s in starts])
    step_s, step_v = [], []
    for k in range(warmup, warmup + steps):
        step_s.append(max(r[k][0] for r in res) if cores >= len(starts) else sum(r[k][0] for r in res) / cores)
        step_v.append(sum(r[k][1] for r in res))
    return step_s, step_v, kind, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    step_s, step_v, kind, cores = cpu_reference_run(args.steps, args.warmup)
    total_s, total_v = sum(step_s), sum(step_v)
    value = total_v / total_s / 1e9
    sample = ("each step = one call of sweepXYZ per source over 1/8 of the 818-FS offsets (slices rotate), "
              "4 sources in 4 processes (the mpi/backup.c source-sharding scheme), state carried across steps")
    line = {
        "impl": "reference", "metric": "GRelax/s", "value": value, "unit": "GRelax/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU reference: (node, offset) visits of serial_new sweepXYZ, "
                   "each visit relaxes both directions of the edge"},
        "cpu_baseline": {"value": value, "unit": "GRelax/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GRelax/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:  # converged sources/s: measured visit rate / (visits per sweep x the reference's recorded sweep counts)
        gold = json.loads((ROOT / "tests" / "golden" / "full_241.json").read_text())
        sweeps = [g["ref_sweeps"] for g in gold if g["label"].startswith("config2")]
        per_sweep = 2_246_171_812
        line["converged_sources_per_s"] = value * 1e9 / (per_sweep * (sum(sweeps) / len(sweeps)))
        line["config"]["sources_per_s_note"] = (f"extrapolated: measured visits/s over {sum(sweeps)/len(sweeps):.0f} sweeps per source "
                                                "(sweep counts recorded from the reference's own converged runs, tests/golden/full_241.json)")
    except Exception:
        pass
    print(json.dumps(line), flush=True)
    return 0


def _ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        t = json.loads((ROOT / "profiles" / "r01_traffic.json").read_text())
        return t["dram_bytes_per_launch"], t["source"]
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import uoparallel_seismic_project_b200 as P
    from uoparallel_seismic_project_b200 import api, dispatch, workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the sweep has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    global DIMS, WORKLOAD
    scaling = "weak"
    if args.workload == "config2":
        v = W.heterogeneous_field(DIMS, 7)
        starts = dispatch.sources_for_rank(rank, world)
    elif args.workload == "config3":
        v = W.heterogeneous_field(DIMS, 7)
        starts = W.starts(111)[dispatch.shard_round_robin(111, rank, world)]
        WORKLOAD, scaling = "241x241x51 heterogeneous slowness (synthetic, seed 7), 818-FS, start-111 sharded over the GPUs", "strong"
    else:
        DIMS = (1201, 1201, 251)
        v = W.heterogeneous_field(DIMS, 11)
        starts = (W.starts(24) * 5)[dispatch.shard_round_robin(24, rank, world)]
        WORKLOAD, scaling = "1201x1201x251 heterogeneous slowness (synthetic, seed 11), 818-FS, 24 sources (start-24 x5) sharded over the GPUs", "strong"
    star = P.make_star(W.star("818"))
    nsrc = len(starts)
    props = torch.cuda.get_device_properties(local)
    sms = props.multi_processor_count
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- device-resident leg: slowness + star + sources already in HBM -------------------------
    stream = torch.cuda.Stream(device=dev)   # not the legacy default stream: the solver captures a CUDA graph on it
    torch.cuda.set_stream(stream)
    ctx = P.SweepContext(device=local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_model(v); ctx.set_star(star); ctx.set_sources(starts)
    for _ in range(args.warmup):
        ctx.run()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    relax = launches = rounds = tiles = 0
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        ev[k][0].record(stream)
        st = ctx.run()
        ev[k][1].record(stream)
        relax += st.relaxations; launches += st.kernel_launches; rounds += st.rounds; tiles += st.tile_visits
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if sampler else None
    tot = dispatch.combine(dist, dev, elapsed_ms=dev_ms, relaxations=relax, sources=nsrc * args.steps,
                           launches=launches)
    assert all(ctx.count_violations(s) == 0 for s in range(nsrc)), "solve did not reach the fixed point"

    # ---- roofline of the dominant kernel (relax_tiled): same steps, every launch event-bracketed --
    pctx = P.SweepContext(device=local, profile_kernels=1)
    pctx.set_stream(stream.cuda_stream)
    pctx.set_model(v); pctx.set_star(star); pctx.set_sources(starts)
    pctx.run()
    k_ms = k_launch = k_relax = 0
    for _ in range(max(1, min(args.steps, 5))):
        flush.zero_()
        st = pctx.run()
        k_ms += st.relax_kernel_ms; k_launch += st.relax_launches; k_relax += st.relaxations
    pctx.close()

    if args.workload == "config4":   # 24 x 1.45 GB of pinned host boxes: the extra workload reports the resident leg only
        ctx.close()
        if rank == 0:
            value = tot["relaxations"] / tot["elapsed_ms"] / 1e6
            peak = sms * 128 * ((clocks or {}).get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
            print(json.dumps({"metric": "GRelax/s", "value": value, "unit": "GRelax/s", "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": tot["elapsed_ms"] / args.steps, "higher_is_better": True,
                              "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": WORKLOAD}, "converged_sources_per_s": tot["sources"] / tot["elapsed_ms"] * 1e3,
                              "clocks": clocks, "gpu_launches": tot["launches"], "e2e": None,
                              "roofline": {"bound": "fp32-issue", "kernel": "relax_tiled<7, fs818>", "unit": "TFLOP/s",
                                           "achieved": 4 * k_relax / (k_ms * 1e-3) / 1e12, "peak": peak,
                                           "frac": 4 * k_relax / (k_ms * 1e-3) / 1e12 / peak}}), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- end-to-end leg: pinned host buffers through the one-shot C-ABI call ----------------------
    hv = torch.from_numpy(v).pin_memory()
    hout = torch.empty((nsrc,) + DIMS, dtype=torch.float32).pin_memory()
    st_arr = api._make_starts(starts)
    ptrs = (ctypes.c_void_p * nsrc)(*[hout[s].data_ptr() for s in range(nsrc)])
    opts = api._opts(device=local)
    for _ in range(args.warmup):
        api.solve_raw(hv.data_ptr(), DIMS, star, st_arr, ptrs, opts)
    barrier()
    e2e_relax = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s2 = api.solve_raw(hv.data_ptr(), DIMS, star, st_arr, ptrs, opts)
        e2e_relax += s2.relaxations
        launches_e2e = s2.kernel_launches
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e = dispatch.combine(dist, dev, elapsed_ms=e2e_ms, relaxations=e2e_relax, sources=nsrc * args.steps, launches=0)
    if rank == 0:
        got = hout[0].numpy()
        assert np.array_equal(got.view(np.uint32), ctx.get_tt(0).view(np.uint32)), "e2e and resident legs disagree"
    ctx.close()
    P.load_library().sweeptt_release_cache()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = tot["relaxations"] / tot["elapsed_ms"] / 1e6          # GRelax/s, whole job
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except OSError:
        pass
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sm_run = (clocks or {}).get("sm_mhz") or sm_max
    lane_ops = 4 * k_relax / (k_ms * 1e-3) / 1e12               # Tlane-op/s of the relax kernel alone
    peak_ops = sms * 128 * sm_max * 1e6 / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tile_bytes = 8 * 8 * 8 * 12                                 # algorithmic bytes per tile visit (8x8x8 nodes x 12 B)
    line = {
        "metric": "GRelax/s", "value": value, "unit": "GRelax/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot["elapsed_ms"] / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sources_per_gpu": nsrc,
                   "loop": ("single persistent launch per solve (work lists built on the device)" if k_launch <= max(1, min(args.steps, 5))
                            else "CUDA-graph WHILE of rounds (device-resident)"),
                   "l2": "flushed between timed steps (512 MiB write, untimed)",
                   "relax_definition": "one pull evaluation tt[n] <- min(tt[n], hd*(v_n+v_m)+tt[m]) with n,m in bounds"},
        "converged_sources_per_s": tot["sources"] / tot["elapsed_ms"] * 1e3,
        "rounds_per_step": rounds / args.steps, "tile_visits_per_step": tiles / args.steps,
        "wall_ms_timed_region": wall_ms,
        "clocks": clocks,
        "gpu_launches": tot["launches"],
        "e2e": {"value": e2e["relaxations"] / e2e["elapsed_ms"] / 1e6, "unit": "GRelax/s",
                "converged_sources_per_s": e2e["sources"] / e2e["elapsed_ms"] * 1e3,
                "ms_per_step": e2e["elapsed_ms"] / args.steps, "timing": "host wall clock around sweeptt_solve()",
                "h2d_bytes_per_step": int(np.prod(DIMS)) * 4, "d2h_bytes_per_step": int(np.prod(DIMS)) * 4 * nsrc},
        "roofline": {
            "bound": "fp32-issue", "kernel": "relax_tiled<7, fs818>",
            "achieved": lane_ops, "peak": peak_ops, "unit": "TFLOP/s", "frac": lane_ops / peak_ops,
            "frac_at_run_clock": lane_ops / (sms * 128 * sm_run * 1e6 / 1e12),
            "grelax_per_s_kernel": k_relax / (k_ms * 1e-3) / 1e9,
            "avg_launch_ms": k_ms / max(1, k_launch), "launches_measured": k_launch,
            "peak_source": f"{sms} SMs x 128 lanes x {sm_max:.0f} MHz (clocks.max.sm), 4 fp32 operations per pull "
                           "(FADD, FMUL, FADD, FMNMX; no FMA allowed by the bit-exactness contract)",
            "hbm": {"achieved": tile_bytes * (tiles / args.steps) / (k_ms / max(1, min(args.steps, 5)) * 1e-3) / 1e9
                    if k_ms else None, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (tile_bytes * (tiles / args.steps) / (k_ms / max(1, min(args.steps, 5)) * 1e-3) / 1e9 / hbm_peak) if k_ms else None,
                    "note": "12 B per node per tile visit; the path is ~50x away from the HBM roof (SURVEY.md 8d)"},
            # dram__bytes_read.sum + dram__bytes_write.sum per launch (= per solve) of the config-2 capture
            "traffic": _ncu_traffic()[0] if args.workload == "config2" else None,
            "traffic_source": _ncu_traffic()[1] if args.workload == "config2" else None,
            "algorithmic_bytes_per_launch": tile_bytes * (tiles / args.steps),
        },
    }
    if not args.no_cpu_baseline and world == 1 and args.workload == "config2":
        try:
            step_s, step_v, kind, cores = cpu_reference_run(2, 1)
            line["cpu_baseline"] = {
                "value": sum(step_v) / sum(step_s) / 1e9, "unit": "GRelax/s", "cores": cores, "kind": kind,
                "sample": "2 timed + 1 warm-up calls of sweepXYZ per source over 1/8 of the 818-FS offsets each, "
                          "4 sources in 4 processes; a CPU visit relaxes both directions of an edge"}
        except Exception as e:  # the baseline must never sink the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "GRelax/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
