"""One native residual-KD training step (crfr_b200.trainer.KDTrainer, eager) for profiling. Usage: python tools/profile_kd_native.py [B] [reps]"""
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L                            # noqa: E402
from crfr_b200.model.resnet import ResNet_34               # noqa: E402
from crfr_b200.trainer import KDTrainer                    # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(7)
nets = [ResNet_34().cuda() for _ in range(3)]
for n in nets:
    for k, p in n.named_parameters():
        if k.endswith("bn2.weight"):
            p.data.fill_(0.5)
teacher, student, assistant = nets
teacher.eval(); student.train(); assistant.train()
tr = KDTrainer(teacher, student, assistant, lr=1e-4, use_graph=False)
x_hr = torch.randn(B, 3, 112, 112, device="cuda")
x_lr = torch.randn(B, 3, 112, 112, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    losses = tr.step(x_hr, x_lr)
    ev[i + 1].record()
torch.cuda.synchronize()
print("losses", [float(l) for l in losses])
print("ms per KD step:", [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)], "launches", L.lib().crfr_launch_count())
