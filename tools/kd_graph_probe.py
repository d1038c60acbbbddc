"""Does CUDA-graph replay help the KD step?  Usage: python tools/kd_graph_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfr_b200.model.resnet import ResNet_34      # noqa: E402
from crfr_b200.trainer import KDTrainer           # noqa: E402

torch.manual_seed(7)
nets = [ResNet_34().cuda() for _ in range(3)]
for n in nets:
    for k, p in n.named_parameters():
        if k.endswith("bn2.weight"):
            p.data.fill_(0.5)
nets[0].eval(); nets[1].train(); nets[2].train()
tr = KDTrainer(*nets, lr=1e-4)
x_hr = torch.randn(256, 3, 112, 112, device="cuda")
x_lr = torch.randn(256, 3, 112, 112, device="cuda")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("eager  %.2f ms" % timed(lambda: tr.step(x_hr, x_lr)))
g = torch.cuda.CUDAGraph()
tr.S.flat_g.zero_(); tr.A.flat_g.zero_()
with torch.cuda.graph(g):
    tr.S.flat_g.zero_(); tr.A.flat_g.zero_()
    tr._native_step(x_hr, x_lr, None)
    tr._optimizer_step(tr.lr)
print("graph  %.2f ms" % timed(g.replay))
