#!/usr/bin/env python
"""bench.py — cell-updates/s of the fused RK4 reaction-diffusion step on a synthetic Turing grid.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 8192]

Workload (BASELINE.json config 5 / north_star): the 2-species Turing model of the reference's
examples/turing.xml written for an n x n grid (spatial_sbml_b200/synthetic.py), n = 8192 per GPU,
dt 0.07.  N > 1 (launched by torchrun, one rank per GPU): the grid is n x (n*N), cut into y-slabs,
halo rows exchanged per RK stage over NCCL → weak scaling.

One "step" = one oneStep() = 4 fused RK stage kernels (the last one also does the RK combine) for
both species on every cell.  `value` is timed with CUDA events on the stream the kernels are
launched on, state resident in HBM; `e2e` goes through the SpatialSimulator binding with pinned
HOST buffers: upload both species, oneStep, download both species, every step.

--impl reference times the reference's own serial CPU implementation (oracle/_ref/ref_driver = its
unmodified sources compiled in place) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

DT = 0.07
S = 2
BYTES_PER_CELL_UPDATE = 104 * S + 4  # 13 fp64 words per species + 4 flag bytes per RK4 step
REF_DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per stage-kernel launch from the committed ncu --set full summary, if any."""
    p = os.path.join(REPO, "profiles", "stage_kernel_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(object):
    """SM clock and throttle reasons sampled through NVML from a thread while the timed region runs
    (an external `nvidia-smi -lms` poller was measured to slow the kernels themselves down)."""

    def __init__(self, index, period=0.01):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = False
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.period = period
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        names = {getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)): "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def run_reference_cpu(n, warmup, steps, model="turing", dt=DT):
    """The reference's serial CPU step on an n x n sample of the workload; returns (cell-updates/s, info)."""
    from spatial_sbml_b200 import synthetic
    if not os.path.exists(REF_DRIVER):
        return None, "oracle/_ref/ref_driver is not built"
    path = tempfile.mktemp(suffix=".xml")
    with open(path, "w") as f:
        f.write(getattr(synthetic, model)(n, n))
    try:
        r = subprocess.run([REF_DRIVER, "time", path, str(n), str(n), "1", repr(dt), str(warmup), str(steps)],
                           capture_output=True, text=True, timeout=1500)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line), None
    except Exception as e:  # noqa: BLE001
        return None, "reference run failed: %r" % (e,)
    finally:
        os.remove(path)


def synthetic_model(name):
    return lambda synthetic: getattr(synthetic, name)


def pick_sample_n(total_steps, budget_s=100.0):
    for n in (1024, 512, 256, 128):
        if total_steps * n * n / 2.2e6 <= budget_s:
            return n
    return 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("BENCH_N", "8192")))
    ap.add_argument("--dim", type=int, default=2, choices=[2, 3], help="3: the n^3 variant of the workload (default n 512), z-slabs")
    ap.add_argument("--model", default="turing", choices=["turing", "brusselator"],
                    help="synthetic model (the headline metric is quoted on turing; brusselator: examples/The_Brusselator.xml kinetics, dt 0.1)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = args.n
    if args.dim == 3 and n == 8192:
        n = 512
    dt = DT if args.dim == 2 else 0.02  # 3-D: six neighbours, D = 4 → explicit stability limit dx^2 / (2 d D) = 0.0208
    make_model = synthetic_model(args.model)
    names = ["species_1", "species_2"] if args.model == "turing" else ["X", "Y"]
    if args.model == "brusselator":
        dt = 0.1 if args.dim == 2 else 0.05
    label = "Turing (examples/turing.xml kinetics)" if args.model == "turing" else "Brusselator (examples/The_Brusselator.xml kinetics)"
    if args.dim == 3:
        config = {"workload": "synthetic %s + z axis, %dx%dx%d cells per GPU, 2 species, fp64, dt %.2f, RK4" % (label, n, n, n, dt),
                  "grid": [n, n, n * world], "cells_per_gpu": n ** 3,
                  "parallelism": "z-slabs x%d, NCCL halo exchange per RK stage" % world if world > 1 else "1 GPU",
                  "l2": "per-step working set %.1f GB per GPU, far larger than the 126 MB L2 (no flush needed)" % (6 * S * n ** 3 * 8 / 1e9)}
    else:
        config = {"workload": "synthetic %s %dx%d cells per GPU, 2 species, fp64, dt %.2f, RK4" % (label, n, n, dt),
                  "grid": [n, n * world], "cells_per_gpu": n * n,
                  "parallelism": "y-slabs x%d, NCCL halo exchange per RK stage" % world if world > 1 else "1 GPU",
                  "l2": "per-step working set %.1f GB per GPU, far larger than the 126 MB L2 (no flush needed)" % (6 * S * n * n * 8 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        total = args.steps + args.warmup
        sn = pick_sample_n(total)
        res, err = run_reference_cpu(sn, args.warmup, args.steps, args.model, dt if args.dim == 2 else (DT if args.model == "turing" else 0.1))
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": err}))
            return
        v = res["cell_updates_per_s"]
        sample = "%dx%d cells of the same synthetic model (2-D), %d timed oneStep calls after %d warm-up, 1 thread (the reference is serial)" % (sn, sn, args.steps, args.warmup)
        print(json.dumps({
            "impl": "reference", "metric": "cell-updates/sec", "value": v, "unit": "cell-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / max(res["steps"], 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config, sample=sample),
            "cpu_baseline": {"value": v, "unit": "cell-updates/s", "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import spatial_sbml_b200 as sb
    from spatial_sbml_b200 import synthetic
    from spatial_sbml_b200.sharded import ShardedSimulator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)

    t_init = time.time()
    if args.dim == 3:
        model = make_model(synthetic)(n, n, n * world)
        shard = ShardedSimulator(model, n, n, n * world, dim=3, device=local_rank)
    else:
        model = make_model(synthetic)(n, n * world)
        shard = ShardedSimulator(model, n, n * world, 1, dim=2, device=local_rank)
    init_s = time.time() - t_init
    cells = shard.num_domain_cells()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    shard.run_steps(0, dt, W)
    # the sampler (NVML init, thread start) costs rank 0 a few ms: set it up BEFORE the barrier, or the other ranks' timed
    # regions start early and then include waiting for rank 0's first halo rows
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sync_all()
    launches0 = shard.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    shard.run_steps(W, dt, K)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = shard.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, float(cells)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms_max, cells_total = float(tm[0]), float(ts[1])
    else:
        ms_max, cells_total = ms, float(cells)
    value = cells_total * K / (ms_max * 1e-3)

    # field sanity
    u = shard.owned(names[0])
    assert np.isfinite(u).all(), "non-finite concentrations after the timed run"

    # ---- end to end through the SpatialSimulator binding with pinned host buffers -------------------
    Ke = max(1, min(args.e2e_steps, K))
    sim = shard.sim
    nzl, nyl, nxl = sim.localShape()
    hbuf = [torch.empty((nzl, nyl, nxl), dtype=torch.float64).pin_memory().numpy() for _ in range(S)]
    for b, name in zip(hbuf, names):
        sim.getVariableCompact(name, out=b)
    e2e = None
    if world == 1:
        # the user-facing call for "state lives on the host": oneStepHost = upload, the four stages and download of
        # successive slabs pipelined on three streams (sb2_step_host); the output of a step is the input of the next
        hout = [torch.empty((nzl, nyl, nxl), dtype=torch.float64).pin_memory().numpy() for _ in range(S)]
        nsl = int(os.environ.get("BENCH_SLABS", "0"))
        sim.oneStepHost(W + K, dt, dict(zip(names, hbuf)), dict(zip(names, hout)), nsl)  # warm-up: streams, events
        hbuf, hout = hout, hbuf
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(Ke):
            sim.oneStepHost(W + K + 1 + i, dt, dict(zip(names, hbuf)), dict(zip(names, hout)), nsl)
            hbuf, hout = hout, hbuf
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1)
        assert np.isfinite(hbuf[0]).all()
        bytes_each_way = S * nzl * nyl * nxl * 8
        e2e = {"value": cells * Ke / (ems * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": bytes_each_way,
               "d2h_bytes_per_step": bytes_each_way, "steps": Ke, "ms_per_step": ems / Ke,
               "api": "SpatialSimulator.oneStepHost (pinned host arrays in and out every step, slab-pipelined)"}
    else:
        # sharded: every rank uploads / downloads its own slab; time = max over ranks
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(Ke):
            for b, name in zip(hbuf, names):
                sim.setVariableCompact(name, b)
            shard.step(W + K + i, dt)
            for b, name in zip(hbuf, names):
                sim.getVariableCompact(name, out=b)
        e1.record()
        sync_all()
        tm = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ems = float(tm[0])
        bytes_each_way = S * nzl * nyl * nxl * 8
        e2e = {"value": cells_total * Ke / (ems * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": bytes_each_way * world,
               "d2h_bytes_per_step": bytes_each_way * world, "steps": Ke, "ms_per_step": ems / Ke}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    specialised = bool(shard.sim.isSpecialized())
    # dominant kernel = the fused stage kernel: 4 launches per step; algorithmic bytes of one step
    # spread over its launches ÷ the average launch duration (the timed region holds nothing else)
    achieved = cells * BYTES_PER_CELL_UPDATE * K / (ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                "kernel": ("rd_jit_first / rd_jit_middle / rd_jit_final (rd_stage kernel compiled for this model's kinetics at set-up: "
                           "TMA-staged rows, straight-line reactions + diffusion stencil, RK combine in stage 4); "
                           if specialised else
                           "rd_stage_kernel (fused RK stage: TMA-staged rows, bytecode reactions + diffusion stencil, RK combine in stage 4); ") +
                          "4 launches per step, figures are the average over the four",
                "specialised": specialised, "specialise_info": shard.sim.getSpecializeInfo(),
                "algorithmic_bytes_per_launch": cells * BYTES_PER_CELL_UPDATE / 4.0,
                "avg_launch_ms": ms / (4.0 * K), "peak_source": peak_src, "rank": 0}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        res, err = run_reference_cpu(512, 3, 80, args.model, dt if args.dim == 2 else (DT if args.model == "turing" else 0.1))
        if res is not None:
            cpu_baseline = {"value": res["cell_updates_per_s"], "unit": "cell-updates/s", "cores": 1, "kind": "reference",
                            "sample": "512x512 cells of the same synthetic model (2-D), 80 timed oneStep calls after 3 warm-up; the reference's own "
                                      "serial code (oracle/_ref), 1 of %d host cores" % (os.cpu_count() or 1),
                            "seconds": res["seconds"], "init_s": res["init_s"]}
        else:
            cpu_baseline = {"value": None, "unit": "cell-updates/s", "cores": 1, "kind": "reference", "sample": err}

    out = {"metric": "cell-updates/sec", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
           "gpu_launches": int(launches), "clocks": clocks, "init_s": init_s, "impl": "ours"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
