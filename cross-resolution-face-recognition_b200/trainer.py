"""Data-parallel FSRNet training loop (the replacement for the loop body at FSR_main.py:219-251).

One process per GPU.  Per step and per rank:
  1. (optional) bicubic 8x upsample of the uint8 LR faces on the GPU (bit-exact with PIL, crfr_bicubic_u8);
  2. the batch is walked in chunks of ``chunk`` images: each chunk runs forward + losses + backward in ONE native call
     (crfr_fsrnet_train_step).  InstanceNorm only -> samples are independent, so chunking is exact; it keeps a chunk's
     2 MiB/img activation maps L2-resident between consecutive kernels and bounds the workspace;
  3. gradients accumulate in one flat fp32 arena laid out in state_dict order, so the three all-reduce buckets
     (decoder | prior+encoder | coarse = reverse execution order) are contiguous slices.  During the LAST chunk's
     backward the native program records a CUDA event as each bucket becomes final; the bucket's NCCL all-reduce is
     issued on a side stream behind that event and overlaps the rest of the backward (only collective in the loop);
  4. fused RMSprop on the flat arena (torch.optim.RMSprop semantics, FSR_main.py:185).

Loss scaling: FSR_main.py:234 divides by 2*train_batch.  With G = global batch (all ranks) and chunk size c, the
chunk's share of the batch-mean losses is c/G, so each chunk is run with loss_div = 2*G*G/c and gradients are
SUM-reduced across ranks: the result equals the single-GPU gradient of the concatenated batch.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops
from .model import FSRnet as M

# state_dict index ranges of the all-reduce buckets, in the order they become ready during backward
BUCKET_PARAM_RANGES = ((167, 202), (33, 167), (0, 33))


def flat_layout(shapes, align=4):
    """Offsets (in elements) of each tensor in a flat arena, every tensor 16-byte aligned."""
    offs, tot = [], 0
    for s in shapes:
        n = 1
        for d in s:
            n *= d
        offs.append(tot)
        tot += (n + align - 1) // align * align
    return offs, tot


def bucket_slices(offs, total, ranges=BUCKET_PARAM_RANGES):
    """Contiguous [start, end) element ranges of the flat arena for each bucket."""
    out = []
    for lo, hi in ranges:
        out.append((offs[lo], offs[hi] if hi < len(offs) else total))
    return out


def chunk_loss_div(global_batch, chunk):
    """loss_div for one chunk so that the summed chunk gradients equal the gradient of FSR_main.py:233-234."""
    return 2.0 * global_batch * global_batch / chunk


class FSRNetTrainer:
    def __init__(self, model, lr=1e-3, alpha=0.99, eps=1e-8, weight_decay=1e-5, chunk=16, w_pix=5.0,
                 engine=L.ENGINE_AUTO, process_group=None, world_size=None):
        import torch.distributed as dist
        self.model = model
        self.lr, self.alpha, self.eps, self.wd = lr, alpha, eps, weight_decay
        self.chunk, self.w_pix, self.engine = chunk, w_pix, engine
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.pg = process_group
        self.world = world_size if world_size is not None else (self.dist.get_world_size(self.pg) if self.dist else 1)
        params = model.ordered_parameters()
        dev = params[0].device
        self.device = dev
        offs, tot = flat_layout([tuple(p.shape) for p in params])
        self.offs, self.total = offs, tot
        self.flat_p = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.flat_sq = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.grad_views = []
        for p, o in zip(params, offs):
            v = self.flat_p[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v                      # parameters now live in the arena (state_dict / optimisers still work)
            self.grad_views.append(self.flat_g[o:o + p.numel()].view(p.shape))
        self.buckets = bucket_slices(offs, tot)
        self.ptable = M._ParamTable([p.data for p in params])
        self.gtable = M._ParamTable(self.grad_views)
        self.losses = torch.zeros((5,), dtype=torch.float32, device=dev)
        self.loss_acc = torch.zeros((5,), dtype=torch.float32, device=dev)
        self.ws = None
        self.outs = None
        if dev.type == "cuda":
            self.comm_stream = torch.cuda.Stream(device=dev)
            self.events = [torch.cuda.Event() for _ in self.buckets]
            for e in self.events:
                e.record()                  # materialise the cudaEvent_t handles
        else:
            self.comm_stream, self.events = None, []

    def reset_optimizer_state(self):
        """The reference re-creates RMSprop every epoch (FSR_main.py:183-185): the square average restarts."""
        self.flat_sq.zero_()

    # -- one native chunk: overridable so that the CPU (gloo) tests can exercise the host logic without a GPU --
    def _native_chunk(self, x, hr, heatmap, labels, outs, loss_div, events):
        b, s = x.shape[0], x.shape[2]
        need = L.lib().crfr_fsrnet_workspace_bytes(b, s, 1)
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        io = M._io(x, outs, (hr, heatmap, labels), loss_div=loss_div, w_pix=self.w_pix)
        for i, e in enumerate(events or ()):
            io.bucket_events[i] = e.cuda_event
        L.call("crfr_fsrnet_train_step", self.engine, self.ptable.arr, self.gtable.arr, C.byref(io),
               self.losses.data_ptr(), self.ws.data_ptr(), self.ws.numel(), ops.stream())

    def _allreduce_bucket(self, k):
        lo, hi = self.buckets[k]
        return self.dist.all_reduce(self.flat_g[lo:hi], op=self.dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def step(self, x, hr, heatmap, labels, lr=None):
        """x [B,3,S,S] fp32 (already upsampled + normalised), hr [B,3,S,S], heatmap [B,S/4,S/4], labels int64
        [B,1,S/4,S/4]; all on this rank's device.  Returns the device tensor (total, L_sr, L_coarse, L_lm, L_ce) of
        this rank's share of the global-batch loss (parts are rank-local batch means)."""
        M.check_input(x) if x.is_cuda else None
        B = x.shape[0]
        G = B * self.world
        self.flat_g.zero_()
        self.loss_acc.zero_()
        if self.outs is None or self.outs[0].shape[0] != min(self.chunk, B) or self.outs[0].shape[2] != x.shape[2]:
            self.outs = M.alloc_outputs(x[:min(self.chunk, B)])
        works = []
        starts = list(range(0, B, self.chunk))
        for ci, s0 in enumerate(starts):
            s1 = min(B, s0 + self.chunk)
            c = s1 - s0
            outs = self.outs if c == self.outs[0].shape[0] else M.alloc_outputs(x[s0:s1])
            last = ci == len(starts) - 1
            use_events = last and self.dist is not None and self.world > 1 and self.events
            self._native_chunk(x[s0:s1], hr[s0:s1], heatmap[s0:s1], labels[s0:s1], outs, chunk_loss_div(G, c),
                               self.events if use_events else None)
            # total adds up over chunks; the parts are chunk means -> weight by c/B
            w = torch.tensor([1.0] + [c / B] * 4, device=self.loss_acc.device)
            self.loss_acc.add_(self.losses * w)
            if use_events:
                main = torch.cuda.current_stream()
                with torch.cuda.stream(self.comm_stream):
                    for k, ev in enumerate(self.events):
                        self.comm_stream.wait_event(ev)
                        works.append(self._allreduce_bucket(k))
                for wk in works:
                    wk.wait()               # main stream waits for the collectives
                del main
        if self.dist is not None and self.world > 1 and not works:
            for k in range(len(self.buckets)):
                self._allreduce_bucket(k).wait()
        self._optimizer_step(self.lr if lr is None else lr)
        return self.loss_acc

    def _optimizer_step(self, lr):
        ops.rmsprop_step(self.flat_p, self.flat_g, self.flat_sq, lr, self.alpha, self.eps, self.wd, gscale=1.0)
