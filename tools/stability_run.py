"""120 training steps of the bench supernet from pinned host input: objective per frame and step time every 10 steps
(checks that the synthetic workload neither diverges nor changes speed as training proceeds)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tdnnf_nas_b200.supernet import Supernet, SupernetConfig
net = Supernet(SupernetConfig(), device=0)
hosts = [net.make_input(i).pin_memory() for i in range(2)]
for i in range(120):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = net.step(hosts[i % 2]); e1.record(); torch.cuda.synchronize()
    if i % 10 == 0 or i > 114: print(i, round(o, 4), round(e0.elapsed_time(e1), 2), flush=True)
net.close()
