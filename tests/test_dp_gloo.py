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
