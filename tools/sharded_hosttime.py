"""torchrun: host-side time per sharded step (enqueue only) vs device time."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch, torch.distributed as dist
from spatial_sbml_b200 import synthetic
from spatial_sbml_b200.sharded import ShardedSimulator
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sh = ShardedSimulator(synthetic.turing(n, n * world), n, n * world, 1, dim=2, device=local)
sh.run_steps(0, 0.07, 5)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
K = 20
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    sh.step(5 + i, 0.07)
e1.record()
t1 = time.time()
torch.cuda.synchronize()
t2 = time.time()
if rank == 0:
    print("host enqueue %.3f ms/step, device %.3f ms/step, wall %.3f ms/step" % (1e3 * (t1 - t0) / K, e0.elapsed_time(e1) / K, 1e3 * (t2 - t0) / K), flush=True)
dist.destroy_process_group()
