#!/usr/bin/env python
"""bench.py — cell-updates/s of the fused RK4 reaction-diffusion step on a synthetic Turing grid.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 8192]

Headline workload (BASELINE.json config 5 / north_star): the 2-species Turing model of the reference's
examples/turing.xml written for an n x n grid (spatial_sbml_b200/synthetic.py), n = 8192 per GPU,
dt 0.07.  N > 1 (launched by torchrun, one rank per GPU): the grid is n x (n*N), cut into y-slabs,
halo rows exchanged per RK stage over NCCL → weak scaling.

One "step" = one oneStep() = 4 fused RK stage kernels (the last one also does the RK combine) for
both species on every cell.  `value` is timed with CUDA events on the stream the kernels are
launched on, state resident in HBM: the timed region of K steps is repeated REPEATS times and the
MEDIAN is reported (min / max beside it).  `e2e` goes through the SpatialSimulator binding with
pinned HOST buffers: upload both species, oneStep, download both species, every step.

The same JSON line carries, under "secondary", the other workloads SURVEY.md §8(d) asks for — the
512^3 z-slab grid (weak), strong scaling at fixed global 8192^2 and 512^3, the synthetic Brusselator
grid and the reference's example models (GPU next to the reference's own CPU step) — each with its
own roofline, and "sharded_check": before anything is timed a small sharded grid is stepped by all
ranks AND by every rank alone on the full grid, and the slabs must agree bit for bit.

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
GOLDEN = os.path.join(REPO, "tests", "golden")


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(dim, n):
    """dram bytes per stage-kernel launch from a committed `ncu --set full` capture of exactly this grid (dim, n per
    GPU), or None: profiles/stage_kernel_traffic.json holds one entry per captured configuration."""
    p = os.path.join(REPO, "profiles", "stage_kernel_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
    except Exception:
        return None
    entries = d.get("entries") if isinstance(d, dict) else None
    if entries is None:  # round-1 layout: one capture, the 2-D 8192^2 grid
        entries = [dict(d, dim=2, n=8192)]
    for e in entries:
        if int(e.get("dim", 0)) == dim and int(e.get("n", 0)) == n:
            return e
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


def run_ref_driver(path, dims, dt, warmup, steps, timeout=1500):
    if not os.path.exists(REF_DRIVER):
        return None, "oracle/_ref/ref_driver is not built"
    try:
        r = subprocess.run([REF_DRIVER, "time", path, str(dims[0]), str(dims[1]), str(dims[2]), repr(dt), str(warmup), str(steps)],
                           capture_output=True, text=True, timeout=timeout)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line), None
    except Exception as e:  # noqa: BLE001
        return None, "reference run failed: %r" % (e,)


def run_reference_cpu(n, warmup, steps, model="turing", dt=DT):
    """The reference's serial CPU step on an n x n sample of the workload; returns (result dict, error)."""
    from spatial_sbml_b200 import synthetic
    path = tempfile.mktemp(suffix=".xml")
    with open(path, "w") as f:
        f.write(getattr(synthetic, model)(n, n))
    try:
        return run_ref_driver(path, (n, n, 1), dt, warmup, steps)
    finally:
        os.remove(path)


def pick_sample_n(total_steps, budget_s=100.0):
    for n in (1024, 512, 256, 128):
        if total_steps * n * n / 2.2e6 <= budget_s:
            return n
    return 64


def numa_bind(local_rank):
    """Run this rank (and therefore allocate its pinned host buffers, first touch) on the NUMA node its GPU hangs off.
    Returns a description of what was done for the JSON line."""
    info = {"cpu_affinity_before": len(os.sched_getaffinity(0))}
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        if hasattr(p, "pci_domain_id"):
            bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        else:
            import pynvml
            pynvml.nvmlInit()
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
            if isinstance(bdf, bytes):
                bdf = bdf.decode()
            bdf = bdf.lower()
            if len(bdf.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs uses 4
                bdf = bdf[4:]
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        info.update({"gpu_bdf": bdf, "numa_node": node, "local_cpulist": cpus})
        want = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                want.update(range(int(a), int(b) + 1))
            elif part:
                want.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = want & allowed
        if use:
            os.sched_setaffinity(0, use)
            info["bound_to_cpus"] = len(use)
        else:
            info["bound_to_cpus"] = 0
            info["note"] = "the GPU's local CPUs are outside this process's cpuset; affinity left unchanged"
        if node >= 0:
            # prefer the GPU's node for every later allocation of this process (set_mempolicy MPOL_PREFERRED = 1)
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))  # x86-64: __NR_set_mempolicy
            info["mempolicy_preferred"] = rc == 0
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)
    return info


class Harness(object):
    """Timing helpers shared by the headline workload and the secondary ones."""

    def __init__(self, torch, dist, rank, local_rank, world, K, W, repeats):
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.K, self.W, self.R = K, W, repeats

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return [float(v) for v in t]

    def time_steps(self, shard, dt, first_step=0, sample_clocks=False, names=None):
        """W warm-up steps, then R timed regions of K steps each, every region bracketed by barrier + synchronize and
        timed with CUDA events on the launching stream; per region the MAX over ranks.  Returns (regions_ms_max_over_ranks,
        this rank's regions_ms, launches per region, next step index)."""
        torch = self.torch
        shard.run_steps(first_step, dt, self.W)
        step0 = first_step + self.W
        mine, launches = [], 0
        # Every region times the SAME K steps: the state after the warm-up is kept on the host and put back before each
        # further region (outside the timed bracket).  The blob-initialised Turing model at dt 0.07 leaves the stable range of
        # explicit RK4 after about 120 steps on a grid this large (the reference does too - the restated step is bit-identical),
        # and non-finite values must not be what is timed.
        snap = dict((nm, shard.sim.getVariableCompact(nm)) for nm in names) if (names and self.R > 1) else None
        # the sampler (NVML init, thread start) costs rank 0 a few ms: set it up BEFORE the first barrier, or the other
        # ranks' timed regions start early and then include waiting for rank 0's first halo rows
        self.clocks = None
        sampler = ClockSampler(self.local_rank) if (sample_clocks and self.rank == 0) else None
        for r in range(self.R):
            if r > 0 and snap is not None:
                for nm in names:
                    shard.sim.setVariableCompact(nm, snap[nm])
            self.sync_all()
            l0 = shard.launch_count()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            shard.run_steps(step0, dt, self.K)
            ev1.record()
            self.sync_all()
            mine.append(ev0.elapsed_time(ev1))
            launches = shard.launch_count() - l0
        step = step0 + self.K
        del snap
        if sampler:
            self.clocks = sampler.stop()
        worst = self.reduce(mine, self.dist.ReduceOp.MAX) if self.world > 1 else list(mine)
        return worst, mine, launches, step


def workload(h, synthetic, ShardedSimulator, model_name, dim, n_per_gpu, scaling, dt, peak, require_specialised=True, sample_clocks=False):
    """Builds the (possibly sharded) simulator of one synthetic workload, times it and returns the result object."""
    import numpy as np
    world = h.world
    if scaling == "weak":
        dims = (n_per_gpu, n_per_gpu, n_per_gpu * world) if dim == 3 else (n_per_gpu, n_per_gpu * world, 1)
    else:  # strong: the global grid is fixed
        dims = (n_per_gpu, n_per_gpu, n_per_gpu) if dim == 3 else (n_per_gpu, n_per_gpu, 1)
    t0 = time.time()
    model = getattr(synthetic, model_name)(*dims[:dim])
    shard = ShardedSimulator(model, dims[0], dims[1], dims[2], dim=dim, device=h.local_rank)
    init_s = time.time() - t0
    specialised = bool(shard.sim.isSpecialized())
    if require_specialised and not specialised:
        raise SystemExit("bench.py: the model-specialised stage kernel is not in use on %s (%s); refusing to report the "
                         "interpreter kernel's numbers as the benchmark" % (dims, shard.sim.getSpecializeInfo()))
    cells = shard.num_domain_cells()
    names = ["species_1", "species_2"] if model_name == "turing" else ["X", "Y"]
    worst, mine, launches, next_step = h.time_steps(shard, dt, sample_clocks=sample_clocks, names=names)
    cells_total = h.reduce([float(cells)], h.dist.ReduceOp.SUM)[0] if world > 1 else float(cells)
    med = statistics.median(worst)
    u = shard.owned(names[0])
    assert np.isfinite(u).all(), "non-finite concentrations after the timed run"
    my_med = statistics.median(mine)
    achieved = cells * BYTES_PER_CELL_UPDATE * h.K / (my_med * 1e-3) / 1e9
    local_n = shard.hi - shard.lo
    res = {"workload": "synthetic %s %s, %s cells in total, %d GPU(s), %s scaling" % (model_name, "x".join(str(d) for d in dims[:dim]), "x".join(str(d) for d in dims[:dim]), world, scaling),
           "grid": list(dims[:dim]), "scaling": scaling, "dt": dt, "value": cells_total * h.K / (med * 1e-3), "unit": "cell-updates/s",
           "ms_per_step": med / h.K, "ms_per_step_min": min(worst) / h.K, "ms_per_step_max": max(worst) / h.K, "repeats": len(worst),
           "cells_total": cells_total, "init_s": init_s, "gpu_launches_per_region": int(launches),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_launch": cells * BYTES_PER_CELL_UPDATE / 4.0, "avg_launch_ms": my_med / (4.0 * h.K),
                        "specialised": specialised, "rank": 0, "rows_or_planes_on_rank0": local_n}}
    return res, shard, next_step, names, (worst, mine, launches)


def sharded_check(h, synthetic, ShardedSimulator, sb):
    """Every rank steps (a) its slab of a small sharded grid and (b) the full grid alone; the slabs must be equal bit for bit
    (there is no reduction inside the step).  Run with the model-specialised kernel (what the timed
    grids use) in 2-D and 3-D, and with boundary conditions that take the unfused schedule (non-zero flux, Dirichlet)."""
    import numpy as np
    torch, dist, world = h.torch, h.dist, h.world
    cases = []
    steps = 8
    old = os.environ.get("SB2_JIT")
    os.environ["SB2_JIT"] = "1"
    try:
        plan = [("turing 2-D", 2, synthetic.turing(512, 64 * world), (512, 64 * world, 1), 0.07, ("species_1", "species_2")),
                ("turing 3-D", 3, synthetic.turing(64, 64, 16 * world), (64, 64, 16 * world), 0.02, ("species_1", "species_2"))]
        if hasattr(synthetic, "turing_bc"):
            plan.append(("turing 2-D, non-zero Neumann flux + Dirichlet side", 2, synthetic.turing_bc(256, 48 * world), (256, 48 * world, 1), 0.05,
                         ("species_1", "species_2")))
        # ... and the step whose state lives on the HOST (what `e2e` times): ShardedSimulator.step_host
        plan.append(("turing 2-D, host arrays in and out every step (step_host)", 2, plan[0][2], plan[0][3], 0.07, ("species_1", "species_2")))
        plan.append(("turing 3-D, host arrays in and out every step (step_host)", 3, plan[1][2], plan[1][3], 0.02, ("species_1", "species_2")))
        for label, dim, model, dims, dt, names in plan:
            ok = True
            if world > 1:
                shard = ShardedSimulator(model, dims[0], dims[1], dims[2], dim=dim, device=h.local_rank)
                if "step_host" in label:
                    cur = [np.ascontiguousarray(shard.sim.getVariableCompact(nm)) for nm in names]
                    nxt = [np.empty_like(c) for c in cur]
                    for t in range(steps):
                        shard.step_host(t, dt, names, cur, nxt)
                        cur, nxt = nxt, cur
                    for nm, c in zip(names, cur):
                        ok = ok and bool(np.array_equal(shard.owned(nm), shard.owned_of(c)))
                else:
                    shard.run_steps(0, dt, steps)
                torch.cuda.synchronize()
                ref = sb.SpatialSimulator(device=h.local_rank)
                ref.initFromString(model, *dims)
                ref.runSteps(0, dt, steps)
                for name in names:
                    whole = ref.getVariableCompact(name)
                    mine = shard.owned(name)
                    want = whole[shard.lo:shard.hi] if dim == 3 else whole[:, shard.lo:shard.hi]
                    ok = ok and bool(np.array_equal(mine, want))
                spec = bool(shard.sim.isSpecialized()) and bool(ref.isSpecialized())
                ref.close()
                shard.sim.close()
                ok = h.reduce([1.0 if ok else 0.0], dist.ReduceOp.MIN)[0] == 1.0
            else:
                spec = None
            cases.append({"case": label, "grid": list(dims[:dim]), "steps": steps, "bit_identical": ok, "specialised_kernel": spec})
    finally:
        if old is None:
            del os.environ["SB2_JIT"]
        else:
            os.environ["SB2_JIT"] = old
    allok = all(c["bit_identical"] for c in cases)
    out = {"bit_identical": allok, "world": world, "cases": cases}
    if world == 1:
        out["note"] = "one rank: nothing is sharded; the check runs at N > 1 (and tests/test_sharded_gpu.py steps slabs on one GPU)"
    return out


def advection_workload(sb, synthetic, torch, peak, K, W, n=4096):
    """CIP-CSLR advection at scale (+40 * d * S_adv bytes per cell-update): two species advected along x and y
    (uniform velocities) with diffusion and a reaction beside it, n x n cells.  Reports the whole step and the direction pass on
    its own, the pass against the HBM roofline with the survey's 40 bytes per cell, pass and species."""
    import ctypes as C
    t0 = time.time()
    sim = sb.SpatialSimulator()
    sim.initFromString(synthetic.advection_uniform(n, n), n, n, 1)
    init_s = time.time() - t0
    dt = 0.05
    cells = sim.getNumDomainCells()
    sim.runSteps(0, dt, W)
    times = []
    for r in range(3):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sim.runSteps(W + r * K, dt, K)
        torch.cuda.synchronize()
        times.append(1e3 * (time.perf_counter() - t1) / K)
    ms_step = statistics.median(times)
    lib = sb.load_library()
    ms_pass, npass = C.c_double(0.0), C.c_int(0)
    rc = lib.sb2_time_advection(C.c_void_p(sim.getDeviceHandle()), dt, 10, C.byref(ms_pass), C.byref(npass))
    u = sim.getVariableCompact("P")
    import numpy as np
    finite = bool(np.isfinite(u).all())
    sim.close()
    S_adv, d = 2, 2
    bytes_pass = 40.0 * cells  # per species and direction pass
    ach = bytes_pass / (ms_pass.value * 1e-3) / 1e9 if ms_pass.value > 0 else None
    bytes_step = cells * (104 * S_adv + 4 + 40 * d * S_adv)
    return {"workload": "synthetic advection (2 species, uniform velocities on both axes, diffusion, one reaction) %dx%d cells, 1 GPU" % (n, n),
            "grid": [n, n], "dt": dt, "cells": int(cells), "init_s": init_s, "ms_per_step": ms_step, "value": cells / (ms_step * 1e-3),
            "unit": "cell-updates/s", "finite": finite, "passes_per_step": npass.value, "rc": rc,
            "roofline_step": {"bound": "hbm", "achieved": bytes_step / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": bytes_step / (ms_step * 1e-3) / 1e9 / peak,
                              "algorithmic_bytes_per_cell_update": 104 * S_adv + 4 + 40 * d * S_adv},
            "roofline": {"bound": "hbm", "kernel": "cip2d_ux_kernel / cip2d_uy_kernel (one fused, warp-independent direction pass of one species with a uniform velocity)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "algorithmic_bytes_per_launch": bytes_pass, "avg_launch_ms": ms_pass.value}}


def example_models(sb, torch, steps=300):
    """The reference's own example models (BASELINE configs 1-4): µs per oneStep on the GPU (model-specialised kernel, what a long
    run settles to) next to the reference's serial CPU step timed in the same run, with the fraction of the HBM roofline."""
    import numpy as np
    rows = []
    peak, _ = measured_peak()
    for case in ("turing_tiny_cl", "brusselator_300", "fish_300", "advection", "transport"):
        path = os.path.join(GOLDEN, "models", case + ".xml")
        if not os.path.exists(path):
            continue
        g = np.load(os.path.join(GOLDEN, case + ".npz"))
        nx, ny, nz = [int(v) for v in g["dims"]]
        dt = float(g["dt"])
        sim = sb.SpatialSimulator()
        sim.initFromFile(path, nx, ny, nz)
        sim.runSteps(0, dt, 20)
        specialised = bool(sim.specializeNow())
        sim.runSteps(20, dt, 20)
        times = []
        first = 40
        l0 = sim.getDeviceLaunchCount()
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            sim.runSteps(first, dt, steps)
            torch.cuda.synchronize()
            times.append(1e6 * (time.perf_counter() - t0) / steps)
            first += steps
        launches = (sim.getDeviceLaunchCount() - l0) / (5.0 * steps)
        cells = sim.getNumDomainCells()
        ns = sim.getNumStagedSpecies()
        sim.close()
        gpu_us = statistics.median(times)
        csteps = max(3, min(200, int(2.0e6 * 2.0 / max(cells, 1))))  # about 2 s of CPU work
        cpu, err = run_ref_driver(path, (nx, ny, nz), dt, 2, csteps, timeout=300)
        cpu_us = 1e6 * cpu["seconds"] / cpu["steps"] if cpu else None
        bytes_step = cells * (104 * ns + 4)
        rows.append({"case": case, "dims": [nx, ny, nz], "cells": int(cells), "species": ns, "dt": dt,
                     "gpu_us_per_step": gpu_us, "gpu_us_per_step_min": min(times), "gpu_us_per_step_max": max(times),
                     "timing": "host clock around %d steps + synchronize, median of 5 (these grids are launch-latency bound)" % steps,
                     "specialised": specialised, "launches_per_step": launches, "gpu_cell_updates_per_s": cells / (gpu_us * 1e-6),
                     "roofline": {"bound": "hbm", "achieved": bytes_step / (gpu_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": bytes_step / (gpu_us * 1e-6) / 1e9 / peak,
                                  "note": "latency-bound: %d KB of state per species, resident in L2" % (cells * 8 // 1024)},
                     "cpu_us_per_step": cpu_us, "cpu_cell_updates_per_s": (cells / (cpu_us * 1e-6)) if cpu_us else None,
                     "cpu": "oracle/_ref/ref_driver (the reference's own code), 1 thread, %d steps" % csteps if cpu else err,
                     "speedup": (cpu_us / gpu_us) if cpu_us else None})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("BENCH_N", "8192")))
    ap.add_argument("--dim", type=int, default=2, choices=[2, 3], help="3: make the n^3 variant (default n 512, z-slabs) the headline workload")
    ap.add_argument("--model", default="turing", choices=["turing", "brusselator"],
                    help="synthetic model (the headline metric is quoted on turing; brusselator: examples/The_Brusselator.xml kinetics, dt 0.1)")
    ap.add_argument("--repeats", type=int, default=int(os.environ.get("BENCH_REPEATS", "5")), help="timed regions of K steps; the median is reported")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="only the headline workload (profiling runs)")
    ap.add_argument("--secondary", default=os.environ.get("BENCH_SECONDARY", "dim3,strong2d,strong3d,brusselator,advection,examples"))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = args.n
    if args.dim == 3 and n == 8192:
        n = 512
    dt = DT if args.dim == 2 else 0.02  # 3-D: six neighbours, D = 4 → explicit stability limit dx^2 / (2 d D) = 0.0208
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
            "config": dict(config, sample=sample, note="extrapolated CPU baseline: the reference's per-cell rate is flat in n and it cannot hold the full grid (> 100 GB of host memory at 8192^2)"),
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
    numa = numa_bind(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)
    h = Harness(torch, dist, rank, local_rank, world, K, W, max(1, args.repeats))
    peak, peak_src = measured_peak()

    # ---- sharded correctness first: nothing is timed on a path that is not bit-identical --------------------------------
    check = sharded_check(h, synthetic, ShardedSimulator, sb)
    if not check["bit_identical"]:
        if rank == 0:
            print(json.dumps({"impl": "ours", "error": "sharded step differs from the single-GPU step", "sharded_check": check}))
        if world > 1:
            dist.destroy_process_group()
        sys.exit(1)

    # ---- headline workload -------------------------------------------------------------------------------------------
    res, shard, next_step, names, (worst, mine, launches) = workload(h, synthetic, ShardedSimulator, args.model, args.dim, n, "weak", dt, peak,
                                                                     sample_clocks=True)
    clocks = h.clocks
    cells = shard.num_domain_cells()
    cells_total = res["cells_total"]
    ms_med = statistics.median(worst)
    value = res["value"]

    # ---- end to end through the SpatialSimulator binding with pinned host buffers -------------------
    Ke = max(1, min(args.e2e_steps, K))
    sim = shard.sim
    nzl, nyl, nxl = sim.localShape()
    hbuf = [torch.empty((nzl, nyl, nxl), dtype=torch.float64).pin_memory().numpy() for _ in range(S)]
    for b, name in zip(hbuf, names):
        sim.getVariableCompact(name, out=b)
    own_rows = shard.hi - shard.lo  # only the rows a rank owns travel
    bytes_each_way = S * (own_rows * nyl * nxl if args.dim == 3 else nzl * own_rows * nxl) * 8
    e2e_regions = []
    if world == 1:
        # the user-facing call for "state lives on the host": oneStepHost = upload, the four stages and download of
        # successive slabs pipelined on three streams (sb2_step_host); the output of a step is the input of the next
        hout = [torch.empty((nzl, nyl, nxl), dtype=torch.float64).pin_memory().numpy() for _ in range(S)]
        nsl = int(os.environ.get("BENCH_SLABS", "0"))
        step = next_step
        h0 = [b.copy() for b in hbuf]  # every region starts from the same host state (see Harness.time_steps)
        sim.oneStepHost(step, dt, dict(zip(names, hbuf)), dict(zip(names, hout)), nsl)  # warm-up: streams, events
        for _ in range(h.R):
            for b, b0 in zip(hbuf, h0):
                np.copyto(b, b0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(Ke):
                step += 1
                sim.oneStepHost(step, dt, dict(zip(names, hbuf)), dict(zip(names, hout)), nsl)
                hbuf, hout = hout, hbuf
            e1.record()
            torch.cuda.synchronize()
            e2e_regions.append(e0.elapsed_time(e1))
        assert np.isfinite(hbuf[0]).all()
        api = "SpatialSimulator.oneStepHost (pinned host arrays in and out every step, slab-pipelined)"
    else:
        # sharded: every rank uploads / downloads its own slab (pinned on the NUMA node of its GPU); time = max over ranks
        step = next_step
        h0 = [b.copy() for b in hbuf]
        hout = [torch.empty((nzl, nyl, nxl), dtype=torch.float64).pin_memory().numpy() for _ in range(S)]
        shard.step_host(step, dt, names, hbuf, hout)  # warm-up: streams, events
        for _ in range(h.R):
            for b, b0 in zip(hbuf, h0):
                np.copyto(b, b0)
            h.sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(Ke):
                step += 1
                shard.step_host(step, dt, names, hbuf, hout)
                hbuf, hout = hout, hbuf
            e1.record()
            h.sync_all()
            e2e_regions.append(e0.elapsed_time(e1))
        e2e_regions = h.reduce(e2e_regions, dist.ReduceOp.MAX)
        assert np.isfinite(shard.owned_of(hbuf[0])).all()
        api = ("per rank: ShardedSimulator.step_host = pinned host arrays in and out every step, upload / four stages / download "
               "pipelined over sub-slabs, one NCCL exchange of four edge rows per step (apron rows recomputed)")
    ems = statistics.median(e2e_regions)
    e2e = {"value": cells_total * Ke / (ems * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": bytes_each_way * world,
           "d2h_bytes_per_step": bytes_each_way * world, "steps": Ke, "repeats": len(e2e_regions), "ms_per_step": ems / Ke,
           "ms_per_step_min": min(e2e_regions) / Ke, "ms_per_step_max": max(e2e_regions) / Ke, "api": api,
           "host_link_GBps_per_rank_each_way": bytes_each_way / (ems / Ke * 1e-3) / 1e9, "numa": numa}
    specialise_info = shard.sim.getSpecializeInfo()
    shard.sim.close()
    del shard, hbuf, h0

    # ---- secondary workloads (SURVEY.md §8d "Multi-GPU reporting", BASELINE.json configs 2-5) --------------------------
    secondary = {}
    want = [] if args.no_secondary else [s for s in args.secondary.split(",") if s]
    def run_secondary(key, *a):
        try:
            r, sh, _, _, _ = workload(h, synthetic, ShardedSimulator, *a)
            sh.sim.close()
            return r
        except SystemExit:
            raise
        except Exception as e:  # noqa: BLE001
            return {"error": repr(e)}
    if "dim3" in want and not (args.dim == 3):
        r = run_secondary("dim3", "turing", 3, 512, "weak", 0.02, peak)
        if "roofline" in r:
            tr = ncu_traffic(3, 512)
            r["roofline"]["traffic"] = tr["dram_bytes_per_launch"] if tr else None
            r["roofline"]["traffic_source"] = tr.get("source") if tr else "no ncu capture of this grid is committed"
        secondary["dim3_weak_512cubed_per_gpu"] = r
    if "strong2d" in want:
        if world > 1:
            secondary["strong_2d_global_8192sq"] = run_secondary("strong2d", "turing", 2, 8192, "strong", DT, peak)
        else:
            secondary["strong_2d_global_8192sq"] = {"same_as": "headline (one GPU: the weak and the strong grid are the same 8192^2 grid)",
                                                    "value": value, "ms_per_step": ms_med / K}
    if "strong3d" in want:
        if world > 1:
            secondary["strong_3d_global_512cubed"] = run_secondary("strong3d", "turing", 3, 512, "strong", 0.02, peak)
        elif "dim3_weak_512cubed_per_gpu" in secondary:
            d3 = secondary["dim3_weak_512cubed_per_gpu"]
            secondary["strong_3d_global_512cubed"] = {"same_as": "dim3_weak_512cubed_per_gpu (one GPU)", "value": d3.get("value"), "ms_per_step": d3.get("ms_per_step")}
    if "brusselator" in want and args.model != "brusselator":
        secondary["brusselator_8192sq_per_gpu"] = run_secondary("brusselator", "brusselator", 2, 8192, "weak", 0.1, peak)

    if "advection" in want and world == 1 and rank == 0:
        try:
            secondary["advection_4096sq"] = advection_workload(sb, synthetic, torch, peak, K, W)
        except Exception as e:  # noqa: BLE001
            secondary["advection_4096sq"] = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if "examples" in want and world == 1:
        try:
            secondary["examples"] = example_models(sb, torch)
        except Exception as e:  # noqa: BLE001
            secondary["examples"] = {"error": repr(e)}

    # dominant kernel = the fused stage kernel: 4 launches per step; algorithmic bytes of one step
    # spread over its launches ÷ the average launch duration (the timed region holds nothing else)
    roofline = res["roofline"]
    tr = ncu_traffic(args.dim, n)
    roofline.update({
        "traffic": tr["dram_bytes_per_launch"] if tr else None,
        "traffic_source": (tr.get("source") or "profiles/stage_kernel_traffic.json") if tr else "no ncu capture of this grid is committed",
        "kernel": "rd_jit_first / rd_jit_middle / rd_jit_final (rd_stage kernel compiled for this model's kinetics at set-up: "
                  "TMA-staged rows, straight-line reactions + diffusion stencil, RK combine in stage 4); "
                  "4 launches per step, figures are the average over the four",
        "specialise_info": specialise_info, "peak_source": peak_src})

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cres, err = run_reference_cpu(512, 3, 80, args.model, dt if args.dim == 2 else (DT if args.model == "turing" else 0.1))
        if cres is not None:
            cpu_baseline = {"value": cres["cell_updates_per_s"], "unit": "cell-updates/s", "cores": 1, "kind": "reference",
                            "sample": "512x512 cells of the same synthetic model (2-D), 80 timed oneStep calls after 3 warm-up; the reference's own "
                                      "serial code (oracle/_ref), 1 of %d host cores; extrapolated to the full grid (the reference's per-cell rate "
                                      "is flat in n, it needs > 100 GB of host memory at 8192^2)" % (os.cpu_count() or 1),
                            "seconds": cres["seconds"], "init_s": cres["init_s"]}
        else:
            cpu_baseline = {"value": None, "unit": "cell-updates/s", "cores": 1, "kind": "reference", "sample": err}

    out = {"metric": "cell-updates/sec", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms_med / K, "ms_per_step_min": min(worst) / K, "ms_per_step_max": max(worst) / K,
           "repeats": len(worst), "timing": "median of %d timed regions of %d steps (CUDA events, max over ranks per region)" % (len(worst), K),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
           "gpu_launches": int(launches) * len(worst), "clocks": clocks, "init_s": res["init_s"], "impl": "ours",
           "sharded_check": check, "secondary": secondary}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
