"""Host-side data-parallel logic on CPU: world_size-2 gloo run of the trainer with the native chunk replaced by a
deterministic stand-in.  Checks the flat-arena layout, the contiguous all-reduce buckets, the chunk/DP loss scaling
and that SUM-reduced gradients equal the single-process gradient of the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_trainer(world):
    from crfr_b200.model.FSRnet import OverallNetwork, weights_init
    from crfr_b200.trainer import FSRNetTrainer

    class FakeTrainer(FSRNetTrainer):
        """The native step differentiates chunk-MEAN losses divided by loss_div.  Stand-in: grad_k = (mean over the
        chunk's samples of mean(x_i)) * (k+1) / loss_div, so with loss_div = 2*G*G/c every sample weighs 1/(2*G*G)
        and the global-batch gradient is the plain sum over all chunks of all ranks."""

        def _native_chunk(self, x, hr, heatmap, labels, outs, loss_div, events):
            s = float(x.double().mean(dim=(1, 2, 3)).mean())
            for k, g in enumerate(self.grad_views):
                g.add_(s * (k + 1) / loss_div)
            c = x.shape[0]
            self.losses.copy_(torch.tensor([s / loss_div, 1.0, 2.0, 3.0, float(c)]))

        def _optimizer_step(self, lr):
            g = self.flat_g + self.wd * self.flat_p
            self.flat_sq.mul_(self.alpha).addcmul_(g, g, value=1 - self.alpha)
            self.flat_p.addcdiv_(g, self.flat_sq.sqrt().add_(self.eps), value=-lr)

    torch.manual_seed(1234)
    net = OverallNetwork()
    net.apply(weights_init)
    return FakeTrainer(net, lr=1e-3, chunk=3, world_size=world), net


def _data(n):
    g = torch.Generator().manual_seed(77)
    return (torch.randn(n, 3, 32, 32, generator=g), torch.randn(n, 3, 32, 32, generator=g),
            torch.rand(n, 8, 8, generator=g), torch.randint(0, 11, (n, 1, 8, 8), generator=g))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr, net = _make_trainer(world)
        x, hr, hm, lbl = _data(8 * world)
        sl = slice(rank * 8, rank * 8 + 8)
        losses = tr.step(x[sl], hr[sl], hm[sl], lbl[sl])
        ret[rank] = (tr.flat_g.clone(), tr.flat_p.clone(), losses.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp2_gradients_equal_single_process_concatenated_batch():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    tr, net = _make_trainer(1)
    x, hr, hm, lbl = _data(8 * world)
    losses = tr.step(x, hr, hm, lbl)
    for r in range(world):
        g, p, l = ret[r]
        assert torch.allclose(g, tr.flat_g, rtol=1e-5, atol=1e-9)
        assert torch.allclose(p, tr.flat_p, rtol=1e-4, atol=1e-6)
    # the per-rank totals add up to the global-batch loss
    assert abs(sum(float(ret[r][2][0]) for r in range(world)) - float(losses[0])) < 1e-6 * abs(float(losses[0])) + 1e-9


def test_flat_layout_and_buckets():
    from crfr_b200.trainer import BUCKET_PARAM_RANGES, bucket_slices, chunk_loss_div, flat_layout
    from oracle import fsrnet_oracle as FO
    shapes = [s for _, s in FO.fsrnet_param_shapes()]
    names = [k for k, _ in FO.fsrnet_param_shapes()]
    offs, tot = flat_layout(shapes)
    assert all(o % 4 == 0 for o in offs) and tot >= 6970311
    b = bucket_slices(offs, tot)
    # reverse execution order: decoder, prior+encoder, coarse; contiguous, disjoint, covering the arena
    assert names[BUCKET_PARAM_RANGES[0][0]].startswith("_fine_sr_decoder.")
    assert names[BUCKET_PARAM_RANGES[1][0]].startswith("_prior_estimation_network.")
    assert names[BUCKET_PARAM_RANGES[1][1] - 1].startswith("_fine_sr_encoder.")
    assert names[BUCKET_PARAM_RANGES[2][1] - 1].startswith("_coarse_sr_network.")
    assert b[2][0] == 0 and b[2][1] == b[1][0] and b[1][1] == b[0][0] and b[0][1] == tot
    # chunk scaling: sum over chunks of c/(2*G*G/c ... ) reproduces 1/(2G) weighting of batch means
    G = 256
    assert abs(sum(c / G / (2.0 * G) for c in (16,) * 16) - sum(1.0 / chunk_loss_div(G, 16) for _ in range(16))) < 1e-12


def test_trainer_moves_parameters_into_arena_and_keeps_state_dict():
    tr, net = _make_trainer(1)
    sd = net.state_dict()
    from oracle import fsrnet_oracle as FO
    ref = FO.build_fsrnet_state_dict(1234)
    assert all(torch.equal(sd[k], ref[k]) for k in ref)
    p0 = net.ordered_parameters()[0]
    assert p0.data.data_ptr() == tr.flat_p.data_ptr()
    tr.flat_p[:4] += 1.0
    assert torch.equal(p0.data.reshape(-1)[:4], tr.flat_p[:4])


# ---- KD trainer (distill_main.py:59-74, 222-225): two flat arenas, one all-reduce bucket each, 1/world scaling ------
def _make_kd_trainer(world):
    from crfr_b200.model.resnet import ResNet_34
    from crfr_b200.trainer import KDTrainer

    class FakeKD(KDTrainer):
        """Stand-in for crfr_kd_train_step: every term of the KD losses is a MEAN over the rank's batch, so the stand-in
        gradient is (batch mean of mean(x_lr_i)) * (k+1) for the student and (mean of x_hr) * (k+2) for the assistant."""

        def _native_step(self, x_hr, x_lr, events):
            x_lr = x_hr if x_lr is None else x_lr
            s, a = float(x_lr.double().mean()), float(x_hr.double().mean())
            for k, g in enumerate(self.S.grad_views):
                g.add_(s * (k + 1))
            for k, g in enumerate(self.A.grad_views):
                g.add_(a * (k + 2))
            self.losses.copy_(torch.tensor([s, a]))

        def _optimizer_step(self, lr):
            for f in (self.S, self.A):
                g = f.flat_g / self.world + self.wd * f.flat_p
                f.flat_sq.mul_(self.alpha).addcmul_(g, g, value=1 - self.alpha)
                f.flat_p.addcdiv_(g, f.flat_sq.sqrt().add_(self.eps), value=-lr)

    nets = []
    for seed in (1, 2, 3):
        torch.manual_seed(seed)
        nets.append(ResNet_34())
    nets[0].eval()
    return FakeKD(nets[0], nets[1], nets[2], lr=1e-3, world_size=world), nets


def _kd_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr, _ = _make_kd_trainer(world)
        g = torch.Generator().manual_seed(5)
        x_hr, x_lr = torch.randn(4 * world, 3, 8, 8, generator=g), torch.randn(4 * world, 3, 8, 8, generator=g)
        sl = slice(rank * 4, rank * 4 + 4)
        losses = tr.step(x_hr[sl], x_lr[sl])
        ret[rank] = (tr.S.flat_g.clone(), tr.A.flat_g.clone(), tr.S.flat_p.clone(), tr.A.flat_p.clone(), losses.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_kd_dp2_update_equals_single_process_concatenated_batch():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_kd_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    tr, _ = _make_kd_trainer(1)
    g = torch.Generator().manual_seed(5)
    x_hr, x_lr = torch.randn(4 * world, 3, 8, 8, generator=g), torch.randn(4 * world, 3, 8, 8, generator=g)
    tr.step(x_hr, x_lr)
    for r in range(world):
        sg, ag, sp, ap, _ = ret[r]
        # SUM over ranks of rank-batch means = world x the global-batch mean; RMSprop divides by world
        assert torch.allclose(sg / world, tr.S.flat_g, rtol=1e-5, atol=1e-9)
        assert torch.allclose(ag / world, tr.A.flat_g, rtol=1e-5, atol=1e-9)
        assert torch.allclose(sp, tr.S.flat_p, rtol=1e-4, atol=1e-6)
        assert torch.allclose(ap, tr.A.flat_p, rtol=1e-4, atol=1e-6)
    assert torch.equal(ret[0][2], ret[1][2])          # replicas stay in lock step


# ---- sharded-gallery identification (utils/eval.py:11 semantics over a row-sharded gallery) ---------------------------
def _torch_topk(p, g, k, index_base=0):
    s = p.float() @ g.float().t()
    v, i = torch.sort(s, dim=1, descending=True, stable=True)       # stable: ties -> lowest index
    return v[:, :k].contiguous(), (i[:, :k] + index_base).to(torch.int32).contiguous()


def _torch_merge(vals, idx, k):
    parts, p, kk = vals.shape
    v = vals.permute(1, 0, 2).reshape(p, parts * kk)
    i = idx.permute(1, 0, 2).reshape(p, parts * kk)
    order = torch.sort(i, dim=1, stable=True)[1]                    # by index first, then a stable sort by score:
    v, i = torch.gather(v, 1, order), torch.gather(i, 1, order)     # equal scores keep the lowest index in front
    o2 = torch.sort(v, dim=1, descending=True, stable=True)[1]
    return torch.gather(v, 1, o2)[:, :k], torch.gather(i, 1, o2)[:, :k]


def _match_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crfr_b200 import ops
        from crfr_b200.utils import utils as U
        ops.cosine_topk = lambda p, g, k, index_base=0, engine=0: _torch_topk(p, g, k, index_base)   # CPU stand-ins for
        ops.topk_merge = _torch_merge                                                                  # the two kernels
        gen = torch.Generator().manual_seed(11)
        gal = torch.nn.functional.normalize(torch.randn(103, 16, generator=gen), dim=1)
        gal[50] = gal[7]                                            # an exact tie across two shards
        pr = torch.nn.functional.normalize(gal[torch.arange(0, 100, 9)] + 0.1 * torch.randn(12, 16, generator=gen), dim=1)
        lo, hi = U.shard_rows(gal.shape[0], world, rank)
        v, i = U.cosine_identify(pr, gal[lo:hi], k=5, normalized=True, index_base=lo, sharded=True)
        ret[rank] = (v.clone(), i.clone(), (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_identification_equals_unsharded():
    world = 3
    ret = mp.Manager().dict()
    mp.spawn(_match_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    gen = torch.Generator().manual_seed(11)
    gal = torch.nn.functional.normalize(torch.randn(103, 16, generator=gen), dim=1)
    gal[50] = gal[7]
    pr = torch.nn.functional.normalize(gal[torch.arange(0, 100, 9)] + 0.1 * torch.randn(12, 16, generator=gen), dim=1)
    v, i = _torch_topk(pr, gal, 5)
    assert [ret[r][2] for r in range(world)] == [(0, 35), (35, 69), (69, 103)]
    for r in range(world):
        assert torch.equal(ret[r][1], i) and torch.allclose(ret[r][0], v)
