"""Debug: train-mode ResNet_34 forward with the BatchNorm statistics from the tile-engine epilogue (option bn_fused_stats)
against the separate statistics pass: first BatchNorm whose (mean, rstd) differ."""
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops                      # noqa: E402
from crfr_b200.model.resnet import ResNet_34               # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(78)
net = ResNet_34().cuda().train()
for k, p in net.named_parameters():
    if k.endswith("bn2.weight"):
        p.data.fill_(0.5)
x = torch.randn(B, 3, 112, 112, device="cuda")
n = L.lib().crfr_resnet34_tape(B, 112, 1, None, 0)
tape = (L.TapeEntry * n)()
L.lib().crfr_resnet34_tape(B, 112, 1, tape, n)
res = {}
for mode in (0, 1):
    ops.set_option("bn_fused_stats", mode)
    outs = net(x)
    ws = outs[0].grad_fn.ws
    torch.cuda.synchronize()
    st = []
    for e in tape:
        if e.kind == 1 and e.stats_off >= 0:
            st.append(ws[e.stats_off:e.stats_off + e.c * 8].view(torch.float32).clone().view(e.c, 2))
    res[mode] = (st, [o.detach().clone() for o in outs])
ops.set_option("bn_fused_stats", 1)
for i, (a, b) in enumerate(zip(*[r[0] for r in (res[0], res[1])])):
    d = (a - b).abs().max().item()
    bad = not torch.isfinite(b).all() or d > 1e-3 * a.abs().max().item()
    if bad or i < 3:
        print("bn %d c=%d: max |diff| %.3e  separate %s  fused %s" % (i, a.shape[0], d, a[:2].flatten().tolist(), b[:2].flatten().tolist()))
    if bad:
        break
print("emb finite:", torch.isfinite(res[1][1][0]).all().item(), " max diff emb:", (res[0][1][0] - res[1][1][0]).abs().max().item())
