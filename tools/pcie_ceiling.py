"""Host-link ceiling of the box (what bounds bench.py's `e2e` at N GPUs): every rank copies pinned host buffers to its GPU, back, and
both ways at once on two streams, all ranks at the same time (barrier before every timed region, CUDA events, the slowest rank
counts).  No kernel of ours runs: these are the numbers a `bandwidthTest`-style tool gives, taken with all N GPUs busy.

  python tools/pcie_ceiling.py                                  (one GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

Prints one JSON line on rank 0.  bench.py's e2e moves `h2d_bytes_per_step` up and `d2h_bytes_per_step` down per step, overlapped:
its per-rank GB/s each way can at best reach `both_ways_GBps_each_way_per_rank` of this tool."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    numa = None
    try:
        import bench
        numa = bench.numa_bind(local)  # the same placement of the pinned buffers as bench.py
    except Exception as e:  # noqa: BLE001
        numa = {"error": repr(e)}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = int(os.environ.get("CEILING_BYTES", str(1 << 30)))
    reps = 5
    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hin.fill_(1)
    hout.fill_(0)
    din = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dout = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    up, down = torch.cuda.Stream(), torch.cuda.Stream()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(do_up, do_down):
        best = []
        for _ in range(reps + 1):
            sync()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            up.wait_event(e0)
            down.wait_event(e0)
            if do_up:
                with torch.cuda.stream(up):
                    din.copy_(hin, non_blocking=True)
            if do_down:
                with torch.cuda.stream(down):
                    hout.copy_(dout, non_blocking=True)
            e1.record(up)
            e2.record(down)
            torch.cuda.current_stream().wait_event(e1)
            torch.cuda.current_stream().wait_event(e2)
            e3 = torch.cuda.Event(enable_timing=True)
            e3.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e3))
        ms = sorted(best[1:])[len(best[1:]) // 2]  # median, first repetition dropped
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    h2d = timed(True, False)
    d2h = timed(False, True)
    both = timed(True, True)
    gb = nbytes / 1e9
    if rank == 0:
        print(json.dumps({
            "tool": "tools/pcie_ceiling.py", "n_gpus": world, "bytes_per_copy": nbytes, "timing": "median of %d, max over ranks" % reps,
            "h2d_GBps_per_rank": gb / (h2d * 1e-3), "d2h_GBps_per_rank": gb / (d2h * 1e-3),
            "both_ways_GBps_each_way_per_rank": gb / (both * 1e-3),
            "aggregate_both_ways_GBps": 2 * world * gb / (both * 1e-3),
            "e2e_floor_ms_per_step_8192sq_2species": 2 * 8192 * 8192 * 8 / 1e9 / (gb / (both * 1e-3)) * 1e3,
            "gpu": torch.cuda.get_device_name(local), "numa_rank0": numa}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
