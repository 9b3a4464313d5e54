"""Run under torchrun (one rank per GPU): the slab-sharded step over NCCL must be BIT-IDENTICAL to the single-GPU step
(no reduction inside the step).  usage: torchrun --nproc-per-node N tools/sharded_check.py [2|3|bc|adv|mt|host2|host3] [steps]
host2 / host3: the steps go through ShardedSimulator.step_host (host arrays in and out every step, one exchange of edge rows per step),
with a plain sharded step in between (which first refreshes the halo rows the host step left unexchanged)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist

from spatial_sbml_b200 import synthetic
from spatial_sbml_b200.sharded import ShardedSimulator
import spatial_sbml_b200 as sb

what = sys.argv[1] if len(sys.argv) > 1 else "2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
names = ("species_1", "species_2")
host = what.startswith("host")
if host:
    what = what[4:]
if what == "2":
    dim, dims = 2, (700, 96 * world + 5, 1)
    model, dt = synthetic.turing(dims[0], dims[1]), 0.07
elif what == "3":
    dim, dims = 3, (200, 40, (20 if host else 6) * world + 3)
    model, dt = synthetic.turing(*dims), 0.02
elif what == "bc":
    dim, dims = 2, (128, 24 * world + 3, 1)
    model, dt = synthetic.turing_bc(dims[0], dims[1]), 0.05
elif what == "mt":  # membrane transport with reactions around it
    dim, dims = 2, (40, 36, 1)
    model, dt = open(os.path.join(REPO, "tests", "golden", "models", "syn_transport_mixed_2d.xml")).read(), 0.05
    names = ("A", "E")
else:  # the reference's advection example
    dim, dims = 2, (101, 101, 1)
    model, dt = open(os.path.join(REPO, "tests", "golden", "models", "advection.xml")).read(), 0.05
    names = ("A_cytosol", "B_cytosol", "C_cytosol")
shard = ShardedSimulator(model, dims[0], dims[1], dims[2], dim=dim, device=local)
if host:
    cur = [np.ascontiguousarray(shard.sim.getVariableCompact(n)) for n in names]
    nxt = [np.empty_like(c) for c in cur]
    for t in range(steps):
        if t == steps // 2:  # a plain sharded step between host steps
            shard.step(t, dt)
            cur = [np.ascontiguousarray(shard.sim.getVariableCompact(n)) for n in names]
            continue
        shard.step_host(t, dt, names, cur, nxt)
        cur, nxt = nxt, cur
else:
    shard.run_steps(0, dt, steps)
torch.cuda.synchronize()
ok = True
ref = sb.SpatialSimulator(device=local)
ref.initFromString(model, *dims)
ref.runSteps(0, dt, steps)
for name in names:
    whole = ref.getVariableCompact(name)
    mine = shard.owned(name)
    if host and not np.array_equal(mine, shard.owned_of(cur[names.index(name)])):
        print("rank %d %s: the host array of the last host step differs from the device state" % (rank, name), flush=True)
        ok = False
    want = whole[shard.lo:shard.hi] if dim == 3 else whole[:, shard.lo:shard.hi]
    same = np.array_equal(mine, want)
    if not same:
        bad = np.argwhere(mine != want)
        rows = sorted(set(int(b[0 if dim == 3 else 1]) for b in bad))
        print("rank %d %s differs: max abs %.3e, %d cells, local owned rows %s ... %s (owned %d rows), x range %d..%d" % (
            rank, name, np.abs(mine - want).max(), len(bad), rows[:6], rows[-3:], mine.shape[0 if dim == 3 else 1],
            bad[:, 2].min(), bad[:, 2].max()), flush=True)
    ok = ok and same
# the stability reduction over the slabs is one MIN all-reduce
ok = ok and shard.stable_time_step(10.0) == ref.getStableTimeStep(10.0)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded_check %s%s world %d steps %d rows [%d,%d) exchanges %d: %s" % ("host " if host else "", what, world, steps, shard.lo, shard.hi, shard.exchanges,
                                                                               "BIT-IDENTICAL" if int(t[0]) else "MISMATCH"), flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t[0]) else 1)
