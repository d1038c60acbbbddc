"""Diagnostic: per-parameter gradient deviation of the native train step from the teacher-forced oracle (bf16 gradient
storage and exact fp32 backward), in state_dict order.  Usage: python tools/forced_diag.py [B] [S]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import forced_check as FC          # noqa: E402
from oracle import fsrnet_oracle as FO         # noqa: E402


class ForcedExact(FO.ForcedPrecision):
    """forced forward, no gradient rounding at all (the exact backward of the stored forward)"""

    def qg(self, x):
        return x


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    net = FC.make_net()
    x, hr, lbl, hm = FO.synthetic_batch(B, S)
    st = FC.Step(net, x.cuda(), (hr.cuda(), hm.cuda(), lbl.cuda().contiguous()))
    losses, grads = st.train_step()
    tape = FC.tape_of(B, S)
    sd = FO.build_fsrnet_state_dict(1234)
    pr = FO.ForcedPrecision(FC.forced_feed(st.ws, tape))
    _, tot, _, g_bf = FO.fsrnet_loss_and_grads(sd, x, hr, hm, lbl, precision=pr)
    pe = ForcedExact(FC.forced_feed(st.ws, tape))
    _, _, _, g_ex = FO.fsrnet_loss_and_grads(sd, x, hr, hm, lbl, precision=pe)
    print("loss", losses[0].item(), tot.item(), "worst layer", max(pr.errors_q))
    rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
    print("%-58s %10s %10s %10s" % ("parameter", "gpu-bf16", "gpu-exact", "bf16-exact"))
    for (k, _), g in zip(net.named_parameters(), grads):
        if FO.fsrnet_dead_param(k) or k in FO.FSRNET_NULL_GRAD:
            continue
        print("%-58s %10.2e %10.2e %10.2e" % (k, rel(g, g_bf[k]), rel(g, g_ex[k]), rel(g_bf[k], g_ex[k])))


if __name__ == "__main__":
    main()
