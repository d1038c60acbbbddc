"""Drop-in for the matcher part of the reference's utils/utils.py plus the cosine identification entry point."""
import numpy as np
import torch

from .. import ops


def l2_norm(x, axis=1):
    """ref: DISTILLATION/model/model_irse.py:16-20.  Returns unit-norm bf16 rows (the matcher's operand format)."""
    assert axis == 1
    return ops.l2norm_bf16(x)


def calculate_accuracy(threshold, dist, actual_issame):
    """ref: utils/utils.py:14-24.  dist / actual_issame may be numpy arrays or tensors; counting runs on the GPU."""
    d = torch.as_tensor(np.asarray(dist) if not torch.is_tensor(dist) else dist).float().cuda()
    s = torch.as_tensor(np.asarray(actual_issame) if not torch.is_tensor(actual_issame) else actual_issame).cuda()
    tp, fp, tn, fn = (int(v) for v in ops.verify_counts(d, s, float(threshold)).tolist())
    tpr = 0 if (tp + fn == 0) else float(tp) / float(tp + fn)
    fpr = 0 if (fp + tn == 0) else float(fp) / float(fp + tn)
    acc = float(tp + tn) / d.numel()
    return tpr, fpr, acc


def calculate_roc(thresholds, embeddings1, embeddings2, actual_issame, nrof_folds=50, pca=0, *, seed=0):
    """ref: utils/utils.py:26-87, same signature (distill_main.py:123-125 passes ``nrof_folds=10, pca=0`` by keyword) plus
    a keyword-only ``seed``: the reference's ``KFold(n_splits, shuffle=True)`` is unseeded, here the folds are those of
    ``KFold(n_splits, shuffle=True, random_state=seed)``.  The dead margin_list work (:47) is dropped.  Pair distances
    run on the GPU and every fold costs two launches: one threshold sweep over its training pairs, one over its test
    pairs (crfr_verify_sweep); thresholds are compared in fp32.  ``pca > 0`` (a per-fold sklearn PCA, never used by the
    reference's own call sites) is not part of the native path."""
    if pca:
        raise NotImplementedError("calculate_roc: pca > 0 is not implemented on the native path")
    e1 = torch.as_tensor(embeddings1).float().cuda()
    e2 = torch.as_tensor(embeddings2).float().cuda()
    assert e1.shape[0] == e2.shape[0] and e1.shape[1] == e2.shape[1]
    nrof_pairs = min(len(actual_issame), e1.shape[0])            # :29
    e1, e2 = e1[:nrof_pairs], e2[:nrof_pairs]
    same = torch.as_tensor(np.asarray(actual_issame)[:nrof_pairs]).cuda()
    thr = torch.as_tensor(np.asarray(thresholds, dtype=np.float32)).cuda()
    dist, _ = ops.pair_verify(e1, e2, 0.0)
    n = dist.numel()
    idx = np.arange(n)
    np.random.RandomState(seed).shuffle(idx)
    sizes = np.full(nrof_folds, n // nrof_folds, int)
    sizes[: n % nrof_folds] += 1
    nt = len(thresholds)
    tprs = np.zeros((nrof_folds, nt)); fprs = np.zeros((nrof_folds, nt))
    accuracy = np.zeros(nrof_folds); best = np.zeros(nrof_folds)

    def rates(counts):
        tp, fp, tn, fn = (counts[:, j].astype(np.float64) for j in range(4))
        with np.errstate(divide="ignore", invalid="ignore"):
            tpr = np.where(tp + fn == 0, 0.0, tp / (tp + fn))
            fpr = np.where(fp + tn == 0, 0.0, fp / (fp + tn))
        return tpr, fpr, (tp + tn) / (tp + fp + tn + fn)

    cur = 0
    for f, sz in enumerate(sizes):
        test = np.sort(idx[cur:cur + sz]); cur += sz
        mask = np.ones(n, bool); mask[test] = False
        tr = torch.from_numpy(np.nonzero(mask)[0].astype(np.int32)).cuda()
        te = torch.from_numpy(test.astype(np.int32)).cuda()
        _, _, acc_train = rates(ops.verify_sweep(dist, same, thr, tr).cpu().numpy())
        tprs[f], fprs[f], acc_test = rates(ops.verify_sweep(dist, same, thr, te).cpu().numpy())
        bi = int(np.argmax(acc_train)); best[f] = thresholds[bi]
        accuracy[f] = acc_test[bi]
    return tprs.mean(0), fprs.mean(0), accuracy.mean(), best


def cosine_identify(probes, gallery, k=5, normalized=False, index_base=0, process_group=None, sharded=False):
    """1:N identification: top-k gallery indices by cosine similarity without materialising the score matrix
    (utils/eval.py:11 semantics: largest first, sorted; ties -> lowest index).  probes [P, D], gallery [G, D] (fp32
    embeddings, or unit-norm bf16 when ``normalized``).

    ``sharded=True`` (one process per GPU, torch.distributed initialised): ``gallery`` is THIS rank's shard of the
    gallery rows, starting at global row ``index_base``; probes are replicated.  Every rank runs the fused GEMM + top-k on
    its shard, the per-rank (score, index) lists - k entries per probe, 8 bytes each: 400 KB per rank at 10 k probes - are
    exchanged with one all-gather each (NCCL), and merged by ``crfr_topk_merge``; every rank returns the global result.
    This is the only exchange step of the matcher."""
    p = probes if normalized else l2_norm(probes)
    g = gallery if normalized else l2_norm(gallery)
    val, idx = ops.cosine_topk(p, g, k, index_base=index_base)
    if not sharded:
        return val, idx
    import torch.distributed as dist
    world = dist.get_world_size(process_group)
    npr = val.shape[0]
    vals = torch.empty((world * npr, k), dtype=val.dtype, device=val.device)      # rank-major concatenation along dim 0
    idxs = torch.empty((world * npr, k), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(vals, val.contiguous(), group=process_group)
    dist.all_gather_into_tensor(idxs, idx.contiguous(), group=process_group)
    return ops.topk_merge(vals.view(world, npr, k), idxs.view(world, npr, k), k)


def shard_rows(total, world, rank):
    """[start, end) of this rank's contiguous share of ``total`` gallery rows (the first total % world ranks get one more)."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
