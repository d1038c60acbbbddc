"""One residual-KD step (teacher eval fwd + student/assistant fwd+bwd) for profiling. Usage: python tools/profile_kd.py [B] [reps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L                            # noqa: E402
from crfr_b200.loss import MSELoss, ResidualKDLoss         # noqa: E402
from crfr_b200.model.resnet import ResNet_34               # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(7)
nets = [ResNet_34().cuda() for _ in range(3)]
for n in nets:
    for k, p in n.named_parameters():
        if k.endswith("bn2.weight"):
            p.data.fill_(0.5)
teacher, student, assistant = nets
teacher.eval(); student.train(); assistant.train()
x = torch.randn(B, 3, 112, 112, device="cuda")
mse, kd = MSELoss(), ResidualKDLoss()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
host = []
ev[0].record()
for i in range(reps):
    t0 = time.perf_counter()
    with torch.no_grad():
        t = teacher(x)
    s, a = student(x), assistant(x)
    l_s = mse(s[0], t[0])
    l_a = sum(kd(t[k], s[k], a[k]) for k in (1, 2, 3, 4)) + kd(t[0], s[0], a[0])
    for n in (student, assistant):
        n.zero_grad(set_to_none=True)
    (l_s + l_a).backward()
    host.append((time.perf_counter() - t0) * 1e3)
    ev[i + 1].record()
torch.cuda.synchronize()
print("losses", l_s.item(), l_a.item())
print("host ms per step:", host)
print("ms per KD step:", [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)], "launches", L.lib().crfr_launch_count())
