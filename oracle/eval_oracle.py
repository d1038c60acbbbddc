"""CPU oracle for the eval matcher (identification top-k, threshold verification).  TEST INFRASTRUCTURE ONLY.

Restates, in numpy / plain torch:
  accuracy             /root/reference/utils/eval.py:6-19 (dup utils/utils.py:156-169), with ``.reshape`` where the
                       reference's ``.view`` fails on modern torch
  calculate_accuracy   /root/reference/utils/utils.py:14-24
  calculate_roc        /root/reference/utils/utils.py:26-87 (K-fold made deterministic with a seed; the reference's
                       KFold(shuffle=True) is unseeded; the O(n^2 log n) margin_list at :47 is dead work)
  l2_norm              /root/reference/DISTILLATION/model/model_irse.py:16-20
  cosine matcher       SURVEY.md 8c-v: scores = l2_norm(p) @ l2_norm(g).T ; topk(maxk, 1, True, True) (eval.py:11);
                       verification "same" <=> dist < thr (utils.py:15) with dist = sum((e1-e2)^2) (:41-43), which on
                       unit vectors is sim > 1 - thr/2.

Pin: ``oracle/make_golden.py`` runs the reference functions (loaded by path) on seeded inputs and stores the results
under ``tests/golden/eval_*.npz``.  Tie order of torch.topk is implementation-defined; this oracle
defines ties as lowest index first and the fixtures are tie-free.
"""
import numpy as np


def l2_norm(x, axis=1):
    return x / np.linalg.norm(x, 2, axis, keepdims=True)


def topk_indices(scores, k):
    """Indices of the k largest entries per row, sorted descending, ties -> lowest index first."""
    order = np.argsort(-scores, axis=1, kind="stable")
    return order[:, :k]


def accuracy(output, target, topk=(1,)):
    output = np.asarray(output)
    target = np.asarray(target).reshape(-1)
    maxk = max(topk)
    pred = topk_indices(output, maxk)                 # [P, maxk]
    correct = pred == target[:, None]
    return [np.float32(correct[:, :k].sum() * np.float32(100.0 / target.shape[0])) for k in topk]


def calculate_accuracy(threshold, dist, actual_issame):
    predict = np.less(dist, threshold)
    same = np.asarray(actual_issame, bool)
    tp = int(np.sum(predict & same)); fp = int(np.sum(predict & ~same))
    tn = int(np.sum(~predict & ~same)); fn = int(np.sum(~predict & same))
    tpr = 0 if tp + fn == 0 else float(tp) / float(tp + fn)
    fpr = 0 if fp + tn == 0 else float(fp) / float(fp + tn)
    return tpr, fpr, float(tp + tn) / dist.size


def pair_sqdist(e1, e2):
    return np.sum(np.square(np.subtract(e1, e2)), axis=1)


def kfold_indices(n, n_splits, seed):
    """sklearn KFold(n_splits, shuffle=True, random_state=seed) split sizes / order, restated."""
    idx = np.arange(n)
    np.random.RandomState(seed).shuffle(idx)
    sizes = np.full(n_splits, n // n_splits, int)
    sizes[: n % n_splits] += 1
    folds, cur = [], 0
    for s in sizes:
        test = np.sort(idx[cur:cur + s]); cur += s
        mask = np.ones(n, bool); mask[test] = False
        folds.append((np.nonzero(mask)[0], test))
    return folds


def calculate_roc(thresholds, e1, e2, actual_issame, nrof_folds=10, seed=0):
    dist = pair_sqdist(e1, e2)
    same = np.asarray(actual_issame, bool)
    nt = len(thresholds)
    tprs = np.zeros((nrof_folds, nt)); fprs = np.zeros((nrof_folds, nt))
    acc = np.zeros(nrof_folds); best = np.zeros(nrof_folds)
    for f, (tr, te) in enumerate(kfold_indices(len(dist), nrof_folds, seed)):
        acc_train = np.array([calculate_accuracy(t, dist[tr], same[tr])[2] for t in thresholds])
        bi = int(np.argmax(acc_train)); best[f] = thresholds[bi]
        for ti, t in enumerate(thresholds):
            tprs[f, ti], fprs[f, ti], _ = calculate_accuracy(t, dist[te], same[te])
        acc[f] = calculate_accuracy(thresholds[bi], dist[te], same[te])[2]
    return tprs.mean(0), fprs.mean(0), acc.mean(), best


def cosine_topk(probes, gallery, k, chunk=1024):
    """fp32 reference matcher on already-normalised (and, for bf16 parity, already bf16-rounded) embeddings."""
    probes = np.asarray(probes, np.float32); gallery = np.asarray(gallery, np.float32)
    idx = np.empty((probes.shape[0], k), np.int64); val = np.empty((probes.shape[0], k), np.float32)
    for s in range(0, probes.shape[0], chunk):
        sc = probes[s:s + chunk] @ gallery.T
        ii = topk_indices(sc, k)
        idx[s:s + chunk] = ii
        val[s:s + chunk] = np.take_along_axis(sc, ii, 1)
    return val, idx


def synthetic_gallery(n_gallery, n_probe, dim=512, noise=0.3, seed=4321):
    """unit-norm gallery, probes = gallery[id] + noise-norm perturbation, renormalised (planted rank-1).
    The perturbation has expected L2 norm ``noise`` (per-element std noise/sqrt(dim)) so rank-1 gaps are >> bf16 error."""
    rng = np.random.RandomState(seed)
    g = l2_norm(rng.standard_normal((n_gallery, dim)).astype(np.float32))
    ids = rng.randint(0, n_gallery, n_probe)
    p = l2_norm(g[ids] + noise * rng.standard_normal((n_probe, dim)).astype(np.float32) / np.float32(np.sqrt(dim)))
    return g.astype(np.float32), p.astype(np.float32), ids
