"""Drop-in for the reference's utils/eval.py: top-k identification accuracy on the GPU."""
import torch

from .. import ops

__all__ = ["accuracy"]


def accuracy(output, target, topk=(1,)):
    """ref: utils/eval.py:6-19 (dup utils/utils.py:156-169).  ``output`` is a [P, G] score matrix; returns a list of
    1-element tensors with the precision@k in percent.  The row-wise top-k runs in ``crfr_topk_rows`` (k <= 8, ties
    resolved to the lowest index; the reference's torch.topk leaves tie order unspecified)."""
    maxk = max(topk)
    batch_size = target.size(0)
    _, pred = ops.topk_rows(output, maxk)
    correct = pred.long().eq(target.view(-1, 1).long())
    res = []
    for k in topk:
        correct_k = correct[:, :k].reshape(-1).float().sum(0, keepdim=True)
        res.append(correct_k.mul_(100.0 / batch_size))
    return res
