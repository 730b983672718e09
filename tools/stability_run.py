"""120 training steps of the bench supernet from pinned host input: objective per frame and step time every 10 steps
(checks that the synthetic workload neither diverges nor changes speed as training proceeds).  Also runs under
torchrun (one process per GPU) to check the data-parallel path the same way."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tdnnf_nas_b200.supernet import Supernet, SupernetConfig

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
pg = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = dist.group.WORLD
net = Supernet(SupernetConfig(xent="--no-xent" not in sys.argv), device=local, rank=rank, world_size=world, process_group=pg)
hosts = [net.make_input(i).pin_memory() for i in range(2)]
steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 120
for i in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = net.step(hosts[i % 2])
    e1.record()
    torch.cuda.synchronize()
    if rank == 0 and (i % 10 == 0 or i > steps - 5 or os.environ.get("VERBOSE")):
        f = net.last_max_change_factors
        print(i, round(o, 4), round(e0.elapsed_time(e1), 2), "min max-change factor %.3g" % float(f.min()),
              "|out| max %.3g" % float(net.head["out"].abs().max()),
              ("xent %.4f" % net.last_xent_objf) if net.cfg.xent else "", flush=True)
net.close()
if world > 1:
    torch.distributed.destroy_process_group()
