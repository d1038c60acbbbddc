// Native ResNet_34 face-embedding network program (student / assistant / teacher of the residual knowledge
// distillation step): the whole forward and backward are sequenced here in C++ on one stream over a caller-provided
// workspace arena, exactly like fsrnet.cu - a tape of fused ops recorded by the forward and replayed in reverse.
//
// Layers: 7x7 s2 stem (lowered im2col + tcgen05 GEMM), 16 BasicBlocks of 3x3 convolutions on the tcgen05 implicit-GEMM
// engine (stage transitions: 3x3 s2 and 1x1 s2 through the lowered recipe 5), train-mode BatchNorm (+ReLU, +residual)
// as the n = 1 grouping of the fused normalise kernels with the running-statistics update, the 25088 -> 512 linear
// head as a tcgen05 GEMM over the NHWC-flattened feature (weights re-ordered from the reference's NCHW flatten), and
// BatchNorm1d on the embedding.
//
// The frozen IR_50 teacher (DISTILLATION/model/model_irse.py) is a second, forward-only program over the same ops:
// BN -> conv3x3 -> PReLU -> conv3x3(stride) -> BN, plus a MaxPool2d(1, stride) / conv1x1+BN shortcut, 24 blocks.
//
// ref: model/resnet.py:18-47 (BasicBlock), :152-225 (ResNet.__init__/_make_layer/forward), :231-236 (ResNet_34);
//      DISTILLATION/model/model_irse.py:49-66 (bottleneck_IR), :103-110 (get_blocks(50)), :129-172 (Backbone).
#include <vector>

#include "common.cuh"
#include "crfr.h"
#include "internal.h"

namespace {

struct Tensor {
  bf16* p = nullptr;
  int n = 0, h = 0, w = 0, c = 0, ld = 0;
  int id = -1;
  float* bn_stats = nullptr;   // train-mode BatchNorm statistics of a convolution output, computed by its epilogue
};
struct Slot {
  bf16* p;
  int ld;
};
enum OpKind { OP_CONV, OP_BN, OP_LINEAR };
struct Op {
  OpKind kind;
  Tensor a, b, out;
  crfr_conv_desc cd;
  int w_idx = -1, b_idx = -1, g_idx = -1, beta_idx = -1;
  int cin_pad = 0;
  float* stats = nullptr;
  bool relu = false, has_res = false, x_needs_grad = true;
};

constexpr int kFeat = 512, kEmb = 512;

// W [o][c*HW + hw] (fp32, reference flatten order) -> Wp [o][hw*C + c] and WpT [hw*C + c][o] (bf16)
__global__ void linear_pack_kernel(const float* __restrict__ w, bf16* __restrict__ wp, bf16* __restrict__ wpt, int O,
                                   int C, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long K = (long long)C * HW;
  if (i >= (long long)O * K) return;
  const int o = (int)(i / K);
  const long long k = i - (long long)o * K;       // hw*C + c
  const int hw = (int)(k / C), c = (int)(k - (long long)hw * C);
  const bf16 v = __float2bfloat16_rn(w[(long long)o * K + (long long)c * HW + hw]);
  if (wp) wp[i] = v;
  if (wpt) wpt[k * O + o] = v;
}

// dW [o][c*HW + hw] += G[hw*C + c][o]
__global__ void linear_unpack_kernel(const float* __restrict__ G, float* __restrict__ dw, int O, int C, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long K = (long long)C * HW;
  if (i >= (long long)O * K) return;
  const int o = (int)(i / K);
  const long long r = i - (long long)o * K;       // c*HW + hw
  const int c = (int)(r / HW), hw = (int)(r - (long long)c * HW);
  dw[i] += G[((long long)hw * C + c) * O + o];
}

// MaxPool2d(kernel 1, stride 2) = pixel subsampling (model_irse.py:53): out[n][y][x][:] = x[n][2y][2x][:]
__global__ void subsample2_kernel(const bf16* __restrict__ x, int x_ld, bf16* __restrict__ out, int out_ld, int oh,
                                  int ow, int groups, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i % groups);
  long long p = i / groups;
  const int ox = (int)(p % ow);
  long long q = p / ow;
  const int oy = (int)(q % oh);
  const long long n = q / oh;
  const long long src = ((n * (2 * oh) + 2 * oy) * (2LL * ow) + 2 * ox) * x_ld + g * 8;
  *reinterpret_cast<uint4*>(out + p * out_ld + g * 8) = *reinterpret_cast<const uint4*>(x + src);
}

// stats[c] = (mean 0, rstd 1): turns the fused normalise kernel into a plain PReLU pass
__global__ void identity_stats_kernel(float* __restrict__ stats, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    stats[2 * i] = 0.f;
    stats[2 * i + 1] = 1.f;
  }
}

struct Net {
  int engine;
  const float* const* params;
  float* const* grads;
  void* const* buffers;
  const crfr_resnet_io* io;
  cudaStream_t st;
  uint8_t* ws;
  size_t ws_bytes, off = 0;
  bool exec;
  int err = CRFR_OK;
  std::vector<Op> tape;
  std::vector<std::vector<Slot>> slots;
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  Tensor feat[4], emb;
  // weight packing: a dry run of the program (exec = false) collects every pack job here, the entry point issues them as
  // one batch, and the real run (packs_done) only re-derives the same arena offsets (as in fsrnet.cu)
  std::vector<crfr_pack_job>* collect = nullptr;
  bool packs_done = false;
  bf16* wp = nullptr;    // packed linear weights (forward)
  bf16* wpt = nullptr;   // transposed (backward)
  int pcur = 0, bncur = 0;

  void* alloc(size_t bytes) {
    size_t a = (off + 1023) & ~(size_t)1023;
    if (a + bytes > ws_bytes) {
      if (err == CRFR_OK) {
        crfr_set_error("resnet34: workspace too small (%zu needed so far, %zu given)", a + bytes, ws_bytes);
        err = CRFR_EWORKSPACE;
      }
      off = a + bytes;
      return ws;
    }
    off = a + bytes;
    return ws + a;
  }
  bool ok() const { return err == CRFR_OK; }
  bool run() const { return exec && err == CRFR_OK; }
  void check(int rc) {
    if (rc != CRFR_OK && err == CRFR_OK) err = rc;
  }
  void check_cuda(cudaError_t e) {
    if (e != cudaSuccess && err == CRFR_OK) {
      crfr_set_error("resnet34: %s", cudaGetErrorString(e));
      err = CRFR_ECUDA;
    }
  }
  Tensor new_tensor(int n, int h, int w, int c) {
    Tensor t;
    t.n = n; t.h = h; t.w = w; t.c = c; t.ld = c < 8 ? 4 : c;
    t.p = (bf16*)alloc((size_t)n * h * w * t.ld * sizeof(bf16));
    t.id = (int)slots.size();
    slots.emplace_back();
    return t;
  }
  float* grad(int idx) { return (grads && idx >= 0) ? grads[idx] : nullptr; }

  // ---- forward ops ----
  Tensor conv(const Tensor& x, int w_idx, int cout, int k, int stride, int pad, bool x_needs_grad) {
    Op op;
    op.kind = OP_CONV;
    crfr_conv_desc& d = op.cd;
    d.n = x.n; d.h = x.h; d.w = x.w; d.cin = x.c; d.cout = cout; d.k = k; d.stride = stride; d.pad = pad;
    d.transposed = 0;
    d.oh = (x.h + 2 * pad - k) / stride + 1;
    d.ow = (x.w + 2 * pad - k) / stride + 1;
    Tensor y = new_tensor(x.n, d.oh, d.ow, cout);
    d.in_ld = x.ld; d.out_ld = y.ld;
    op.cin_pad = x.c < 8 ? 4 : x.c;
    const int T = k * k;
    void* wpk = alloc((size_t)T * cout * op.cin_pad * sizeof(bf16));
    if (collect && ok()) collect->push_back({params[w_idx], wpk, T, cout, x.c, op.cin_pad, (long long)x.c * T, T, 0, 0});
    // every convolution of this network feeds a BatchNorm: in training its batch statistics come out of the epilogue
    float* bn_stats = io->training ? (float*)alloc((size_t)cout * 2 * sizeof(float)) : nullptr;
    y.bn_stats = bn_stats;
    if (run()) {
      if (!packs_done) check(crfr_pack_weight(params[w_idx], wpk, T, cout, x.c, op.cin_pad, (long long)x.c * T, T, 1, st));
      if (bn_stats)
        check(crfr_conv_fwd_bnstats(engine, &d, x.p, wpk, op.cin_pad, y.p, bn_stats, io->eps, scratch, scratch_bytes, st));
      else
        check(crfr_conv_fwd(engine, &d, x.p, wpk, op.cin_pad, nullptr, y.p, nullptr, nullptr, io->eps, scratch,
                            scratch_bytes, st));
    }
    op.a = x; op.out = y; op.w_idx = w_idx; op.x_needs_grad = x_needs_grad;
    tape.push_back(op);
    return y;
  }

  // BatchNorm (batch statistics over n*h*w in training, running statistics otherwise) + optional residual + ReLU
  Tensor bn(const Tensor& y, int g_idx, int relu, const Tensor* res, int alpha_idx = -1, int bn_index = -1) {
    const int bi = bn_index >= 0 ? bn_index : bncur++;
    const long long count = (long long)y.n * y.h * y.w;
    Tensor out = new_tensor(y.n, y.h, y.w, y.c);
    const bool have_stats = io->training && y.bn_stats;   // computed by the producing convolution's epilogue
    float* stats = have_stats ? y.bn_stats : (float*)alloc((size_t)y.c * 2 * sizeof(float));
    float* rmean = buffers ? (float*)buffers[3 * bi] : nullptr;
    float* rvar = buffers ? (float*)buffers[3 * bi + 1] : nullptr;
    long long* nbt = buffers ? (long long*)buffers[3 * bi + 2] : nullptr;
    if (run()) {
      if (io->training) {
        if (!have_stats) check(crfr_norm_stats(y.p, 1, (int)count, y.c, y.ld, io->eps, stats, scratch, scratch_bytes, st));
        if (rmean && rvar)
          check(crfr_bn_update_running(stats, rmean, rvar, nbt, y.c, count, io->momentum, io->eps, st));
      } else {
        check(crfr_bn_running_to_stats(rmean, rvar, y.c, io->eps, stats, st));
      }
      check(crfr_norm_act_fwd(y.p, y.ld, stats, params[g_idx], params[g_idx + 1],
                              alpha_idx >= 0 ? params[alpha_idx] : nullptr, relu, res ? res->p : nullptr,
                              res ? res->ld : 8, out.p, out.ld, 1, (int)count, y.c, st));
    }
    Op op;
    op.kind = OP_BN;
    op.a = y; op.out = out; op.stats = stats; op.g_idx = g_idx; op.beta_idx = g_idx + 1; op.relu = relu != 0;
    op.has_res = res != nullptr;
    if (res) op.b = *res;
    tape.push_back(op);
    return out;
  }

  Tensor block(const Tensor& x, int planes, int stride, bool down) {   // BasicBlock, model/resnet.py:18-47
    const int p0 = pcur;
    pcur += down ? 9 : 6;
    Tensor y1 = conv(x, p0 + 0, planes, 3, stride, 1, true);
    Tensor a1 = bn(y1, p0 + 1, 1, nullptr);
    Tensor y2 = conv(a1, p0 + 3, planes, 3, 1, 1, true);
    // module order of the BatchNorms is bn1, bn2, downsample.1: reserve bn2's slot before the downsample branch
    const int bn2 = bncur++;
    Tensor res = x;
    if (down) {
      Tensor yd = conv(x, p0 + 6, planes, 1, stride, 0, true);
      res = bn(yd, p0 + 7, 0, nullptr);
    }
    const int save = bncur;
    bncur = bn2;
    Tensor out = bn(y2, p0 + 4, 1, &res);
    bncur = save;
    return out;
  }

  Tensor linear(const Tensor& a, int w_index = -1) {
    const int B = a.n, K = a.h * a.w * a.c;
    Tensor y = new_tensor(B, 1, 1, kEmb);
    wp = (bf16*)alloc((size_t)kEmb * K * sizeof(bf16));
    wpt = io->training ? (bf16*)alloc((size_t)kEmb * K * sizeof(bf16)) : nullptr;
    const int w_idx = w_index >= 0 ? w_index : pcur;
    if (w_index < 0) pcur += 2;
    if (run()) {
      const long long total = (long long)kEmb * K;
      linear_pack_kernel<<<crfr_cdiv(total, 256), 256, 0, st>>>(params[w_idx], wp, wpt, kEmb, a.c, a.h * a.w);
      CRFR_COUNT_LAUNCH();
      check_cuda(cudaGetLastError());
      TcGemm g{a.p, B, 1, 1, K, K, wp, 1, 0, 1, kEmb, 64, y.p, kEmb, 0, params[w_idx + 1]};
      check(crfr_tc_gemm(g, st));
    }
    Op op;
    op.kind = OP_LINEAR;
    op.a = a; op.out = y; op.w_idx = w_idx; op.b_idx = w_idx + 1;
    tape.push_back(op);
    return y;
  }

  void to_nchw(const Tensor& t, float* dst) {
    if (dst && run()) check(crfr_nhwc_bf16_to_nchw_f32(t.p, dst, t.n, t.c, t.h, t.w, t.ld, st));
  }

  size_t scratch_need(int B, int S) const {
    size_t m = crfr_norm_ws_bytes(1, B * (S / 2) * (S / 2), 64);
    const int q = S / 2;
    const crfr_conv_desc shapes[] = {
        {B, S, S, 3, 64, 7, 2, 3, q, q, 4, 64, 0},
        {B, q, q, 64, 64, 3, 1, 1, q, q, 64, 64, 0},
        {B, q, q, 64, 128, 3, 2, 1, q / 2, q / 2, 64, 128, 0},
        {B, q / 2, q / 2, 128, 128, 3, 1, 1, q / 2, q / 2, 128, 128, 0},
        {B, q / 2, q / 2, 128, 256, 3, 2, 1, q / 4, q / 4, 128, 256, 0},
        {B, q / 4, q / 4, 256, 256, 3, 1, 1, q / 4, q / 4, 256, 256, 0},
        {B, q / 4, q / 4, 256, 512, 3, 2, 1, q / 8, q / 8, 256, 512, 0},
        {B, q / 8, q / 8, 512, 512, 3, 1, 1, q / 8, q / 8, 512, 512, 0}};
    for (const crfr_conv_desc& d : shapes) {
      const size_t b = crfr_conv_workspace_bytes(&d);
      if (b > m) m = b;
    }
    const size_t lin = sizeof(float) * (size_t)kEmb * kFeat * (q / 8) * (q / 8) + 4096;   // linear wgrad accumulator
    return (m > lin ? m : lin) + 4096;
  }

  void forward() {
    const int B = io->batch, S = io->size;
    pcur = 0; bncur = 0;
    scratch_bytes = scratch_need(B, S);
    scratch = alloc(scratch_bytes);
    Tensor x4 = new_tensor(B, S, S, 3);
    if (run()) check(crfr_nchw_f32_to_nhwc_bf16(io->x, x4.p, B, 3, S, S, 4, 4, st));
    // stem (model/resnet.py:208-211; the max-pool is commented out in the reference)
    Tensor y = conv(x4, 0, 64, 7, 2, 3, false);
    pcur = 1;
    Tensor a = bn(y, 1, 1, nullptr);
    pcur = 3;
    const int layers[4] = {3, 4, 6, 3}, planes[4] = {64, 128, 256, 512};
    for (int l = 0; l < 4; ++l) {
      for (int b = 0; b < layers[l]; ++b) a = block(a, planes[l], (l > 0 && b == 0) ? 2 : 1, l > 0 && b == 0);
      feat[l] = a;
      to_nchw(a, io->feat[l]);
    }
    Tensor o1 = bn(a, pcur, 0, nullptr);      // bn_o1
    pcur += 2;
    Tensor yfc = linear(o1);                  // fc (pcur += 2 inside)
    emb = bn(yfc, pcur, 0, nullptr);          // bn_o2 (BatchNorm1d)
    pcur += 2;
    to_nchw(emb, io->emb);
  }

  // ---- IR_50 (forward only, eval-mode BatchNorm) ----
  Tensor prelu(const Tensor& y, int alpha_idx, float* ident) {   // plain PReLU pass
    Tensor out = new_tensor(y.n, y.h, y.w, y.c);
    if (run())
      check(crfr_norm_act_fwd(y.p, y.ld, ident, nullptr, nullptr, params[alpha_idx], 0, nullptr, 8, out.p, out.ld, 1,
                              y.n * y.h * y.w, y.c, st));
    return out;
  }
  Tensor subsample2(const Tensor& x) {
    Tensor out = new_tensor(x.n, x.h / 2, x.w / 2, x.c);
    if (run()) {
      const long long total = (long long)out.n * out.h * out.w * (x.c / 8);
      subsample2_kernel<<<crfr_cdiv(total, 256), 256, 0, st>>>(x.p, x.ld, out.p, out.ld, out.h, out.w, x.c / 8, total);
      CRFR_COUNT_LAUNCH();
      check_cuda(cudaGetLastError());
    }
    return out;
  }

  // parameter order (named_parameters): input_layer (4), output_layer (6), body blocks (7, or 10 with a conv shortcut);
  // BatchNorm order (named_buffers): input_layer.1, output_layer.0, output_layer.4, then per block
  // [shortcut_layer.1], res_layer.0, res_layer.4
  void forward_ir50() {
    const int B = io->batch, S = io->size;
    scratch_bytes = scratch_need_ir50(B, S);
    scratch = alloc(scratch_bytes);
    float* ident = (float*)alloc(512 * 2 * sizeof(float));
    if (run()) {
      identity_stats_kernel<<<4, 128, 0, st>>>(ident, 512);
      CRFR_COUNT_LAUNCH();
    }
    Tensor x4 = new_tensor(B, S, S, 3);
    if (run()) check(crfr_nchw_f32_to_nhwc_bf16(io->x, x4.p, B, 3, S, S, 4, 4, st));
    Tensor a = bn(conv(x4, 0, 64, 3, 1, 1, false), 1, 0, nullptr, 3, 0);     // input_layer: conv, BN, PReLU
    int p = 10, bi = 3;
    const int units[4] = {3, 4, 14, 3}, depth[4] = {64, 128, 256, 512};
    int in_ch = 64;
    for (int l = 0; l < 4; ++l)
      for (int u = 0; u < units[l]; ++u) {   // bottleneck_IR, model_irse.py:49-66
        const int stride = u == 0 ? 2 : 1, d = depth[l];
        const bool conv_sc = in_ch != d;
        Tensor sc = a;
        if (conv_sc) {
          sc = bn(conv(a, p, d, 1, stride, 0, false), p + 1, 0, nullptr, -1, bi);
          p += 3; bi += 1;
        } else if (stride == 2) {
          sc = subsample2(a);
        }
        Tensor r = bn(a, p, 0, nullptr, -1, bi);                       // res_layer.0
        r = prelu(conv(r, p + 2, d, 3, 1, 1, false), p + 3, ident);    // res_layer.1, .2
        r = conv(r, p + 4, d, 3, stride, 1, false);                    // res_layer.3
        a = bn(r, p + 5, 0, &sc, -1, bi + 1);                          // res_layer.4 + shortcut
        p += 7; bi += 2;
        in_ch = d;
        if (u == units[l] - 1) {   // stage output: 64@56^2, 128@28^2, 256@14^2, 512@7^2 - the shapes of ResNet_34's x1..x4,
          feat[l] = a;             // i.e. the t_k of the residual-KD loss when IR_50 is the teacher (distill_main.py:59,68-69)
          to_nchw(a, io->feat[l]);
        }
      }
    Tensor o = bn(a, 4, 0, nullptr, -1, 1);                            // output_layer.0 (Dropout: identity in eval)
    Tensor y = linear(o, 6);                                           // output_layer.3
    emb = bn(y, 8, 0, nullptr, -1, 2);                                 // output_layer.4 (BatchNorm1d)
    to_nchw(emb, io->emb);
  }

  size_t scratch_need_ir50(int B, int S) const {
    size_t m = crfr_norm_ws_bytes(1, B * S * S, 64);
    const int q = S / 2;
    const crfr_conv_desc shapes[] = {
        {B, S, S, 3, 64, 3, 1, 1, S, S, 4, 64, 0},
        {B, S, S, 64, 64, 3, 2, 1, q, q, 64, 64, 0},
        {B, q, q, 64, 128, 3, 2, 1, q / 2, q / 2, 64, 128, 0},
        {B, q, q, 64, 128, 1, 2, 0, q / 2, q / 2, 64, 128, 0},
        {B, q / 2, q / 2, 128, 256, 3, 2, 1, q / 4, q / 4, 128, 256, 0},
        {B, q / 4, q / 4, 256, 512, 3, 2, 1, q / 8, q / 8, 256, 512, 0},
        {B, q / 8, q / 8, 512, 512, 3, 1, 1, q / 8, q / 8, 512, 512, 0}};
    for (const crfr_conv_desc& d : shapes) {
      const size_t b = crfr_conv_workspace_bytes(&d);
      if (b > m) m = b;
    }
    return m + 4096;
  }

  // ---- backward ----
  void add_slot(const Tensor& t, bf16* p, int ld) { slots[t.id].push_back({p, ld}); }
  bool has_grad(const Tensor& t) const { return t.id >= 0 && !slots[t.id].empty(); }
  void squash(const Tensor& t, size_t max_slots) {
    std::vector<Slot>& s = slots[t.id];
    while (s.size() > max_slots) {
      const bool three = s.size() >= 3;
      bf16* out = (bf16*)alloc((size_t)t.n * t.h * t.w * t.ld * sizeof(bf16));
      const size_t k = s.size();
      Slot a = s[k - 1], b = s[k - 2], c = three ? s[k - 3] : Slot{nullptr, t.ld};
      if (run()) check(crfr_add_n(a.p, a.ld, b.p, b.ld, c.p, c.ld, out, t.ld, (long long)t.n * t.h * t.w, t.ld, st));
      s.resize(k - (three ? 3 : 2));
      s.push_back({out, t.ld});
    }
  }
  void seed_grad(const Tensor& t, const float* g) {   // external gradient of an output (fp32 NCHW)
    if (!g) return;
    bf16* d = (bf16*)alloc((size_t)t.n * t.h * t.w * t.ld * sizeof(bf16));
    if (run()) check(crfr_nchw_f32_to_nhwc_bf16(g, d, t.n, t.c, t.h, t.w, t.ld, t.ld, st));
    add_slot(t, d, t.ld);
  }

  void backward() {
    for (int i = (int)tape.size() - 1; i >= 0 && ok(); --i) {
      Op& op = tape[i];
      if (!has_grad(op.out)) continue;
      switch (op.kind) {
        case OP_BN: {
          squash(op.out, 2);
          std::vector<Slot>& s = slots[op.out.id];
          const Tensor& y = op.a;
          const long long count = (long long)y.n * y.h * y.w;
          const size_t bytes = (size_t)count * y.ld * sizeof(bf16);
          bf16* dz = (op.has_res || s.size() > 1) ? (bf16*)alloc(bytes) : nullptr;   // only if somebody consumes it
          bf16* dy = (bf16*)alloc(bytes);
          if (run())
            check(crfr_norm_act_bwd(s[0].p, s[0].ld, s.size() > 1 ? s[1].p : nullptr, s.size() > 1 ? s[1].ld : 8, y.p, y.ld,
                                    op.stats, params[op.g_idx], params[op.beta_idx], nullptr, op.relu ? 1 : 0,
                                    op.has_res ? op.b.p : nullptr, op.has_res ? op.b.ld : 8, dz, y.ld, dy, y.ld,
                                    grad(op.g_idx), grad(op.beta_idx), nullptr, 1, (int)count, y.c, scratch, scratch_bytes,
                                    st));
          add_slot(y, dy, y.ld);
          if (op.has_res) add_slot(op.b, dz, y.ld);
          break;
        }
        case OP_CONV: {
          squash(op.out, 1);
          Slot dy = slots[op.out.id][0];
          crfr_conv_desc d = op.cd;
          d.out_ld = dy.ld;
          if (run()) check(crfr_conv_wgrad(engine, &d, op.a.p, dy.p, grad(op.w_idx), nullptr, scratch, scratch_bytes, st));
          if (op.x_needs_grad) {
            const Tensor& x = op.a;
            bf16* dx = (bf16*)alloc((size_t)x.n * x.h * x.w * x.ld * sizeof(bf16));
            const int T = d.k * d.k;
            void* wt = alloc((size_t)T * d.cin * d.cout * sizeof(bf16));
            if (collect && ok()) collect->push_back({params[op.w_idx], wt, T, d.cin, d.cout, d.cout, T, (long long)d.cin * T, 0, 0});
            if (run()) {
              if (!packs_done)
                check(crfr_pack_weight(params[op.w_idx], wt, T, d.cin, d.cout, d.cout, T, (long long)d.cin * T, 1, st));
              check(crfr_conv_dgrad(engine, &d, dy.p, wt, d.cout, dx, scratch, scratch_bytes, st));
            }
            add_slot(x, dx, x.ld);
          }
          break;
        }
        case OP_LINEAR: {
          squash(op.out, 1);
          Slot dy = slots[op.out.id][0];
          const Tensor& a = op.a;
          const int B = a.n, K = a.h * a.w * a.c;
          bf16* da = (bf16*)alloc((size_t)B * K * sizeof(bf16));
          if (run()) {
            TcGemm g{dy.p, B, 1, 1, kEmb, dy.ld, wpt, 1, 0, 1, K, 0, da, K, 0, nullptr};
            check(crfr_tc_gemm(g, st));
            float* G = (float*)scratch;
            check_cuda(cudaMemsetAsync(G, 0, sizeof(float) * (size_t)K * kEmb, st));
            TcWgrad wg{a.p, B, 1, 1, K, K, dy.p, kEmb, dy.ld, 0, G};
            check(crfr_tc_wgrad_raw(wg, st));
            if (grad(op.w_idx)) {
              const long long total = (long long)kEmb * K;
              linear_unpack_kernel<<<crfr_cdiv(total, 256), 256, 0, st>>>(G, grad(op.w_idx), kEmb, a.c, a.h * a.w);
              CRFR_COUNT_LAUNCH();
              check_cuda(cudaGetLastError());
            }
            if (grad(op.b_idx)) check(crfr_colsum(dy.p, dy.ld, kEmb, B, grad(op.b_idx), st));
          }
          add_slot(a, da, K / (a.h * a.w));
          break;
        }
      }
    }
  }
};

int check_io(const crfr_resnet_io* io, const char* who) {
  CRFR_CHECK_ARG(io && io->batch > 0 && io->size == 112, "%s: batch/size invalid (input must be [B,3,112,112])", who);
  CRFR_CHECK_ARG(io->x && io->emb, "%s: null tensor pointer", who);
  CRFR_CHECK_ARG(io->eps > 0.f && io->momentum >= 0.f && io->momentum <= 1.f, "%s: eps/momentum invalid", who);
  return CRFR_OK;
}

void init_net(Net& net, int engine, const float* const* params, float* const* grads, void* const* buffers,
              const crfr_resnet_io* io, void* ws, size_t ws_bytes, cudaStream_t st, bool exec) {
  net.engine = engine; net.params = params; net.grads = grads; net.buffers = buffers; net.io = io; net.st = st;
  net.ws = (uint8_t*)ws; net.ws_bytes = ws_bytes; net.exec = exec;
}

}  // namespace

extern "C" size_t crfr_resnet34_workspace_bytes(int batch, int size, int training) {
  if (batch <= 0 || size != 112) return 0;
  crfr_resnet_io io = {};
  io.batch = batch; io.size = size; io.training = training; io.momentum = 0.1f; io.eps = 1e-5f;
  static float dummy = 0.f;   // non-null marker so that the dry run sizes every optional buffer
  Net net;
  init_net(net, CRFR_ENGINE_AUTO, nullptr, nullptr, nullptr, &io, nullptr, ~(size_t)0 >> 2, nullptr, false);
  net.forward();
  if (training) {
    for (int l = 0; l < 4; ++l) net.seed_grad(net.feat[l], &dummy);
    net.seed_grad(net.emb, &dummy);
    net.backward();
  }
  return net.off + 65536;
}

// Layout of the stored forward tensors of the ResNet_34 program in the workspace (a dry run: nothing is launched): one
// entry per recorded op in program order - kind 0 convolution, 1 BatchNorm (+ residual) (+ ReLU), 7 linear head; out_off /
// in_off / stats_off are bytes from the workspace base.  The parity tests read the stored bf16 activations through this
// table and replay them into the CPU oracle (teacher-forced forward / backward checks, tests/test_resnet_forced_gpu.py).
extern "C" int crfr_resnet34_tape(int batch, int size, int training, crfr_tape_entry* entries, int max_entries) {
  if (batch <= 0 || size != 112) return -1;
  crfr_resnet_io io = {};
  io.batch = batch; io.size = size; io.training = training; io.momentum = 0.1f; io.eps = 1e-5f;
  Net net;
  init_net(net, CRFR_ENGINE_AUTO, nullptr, nullptr, nullptr, &io, nullptr, ~(size_t)0 >> 2, nullptr, false);
  net.forward();
  const int count = (int)net.tape.size();
  for (int i = 0; i < count && i < max_entries && entries; ++i) {
    const Op& op = net.tape[i];
    crfr_tape_entry& e = entries[i];
    const Tensor& t = op.out;
    e.kind = op.kind == OP_CONV ? 0 : (op.kind == OP_BN ? 1 : 7);
    e.n = t.n; e.h = t.h; e.w = t.w; e.c = t.c; e.ld = t.ld;
    e.out_off = t.p ? (long long)((uint8_t*)t.p - (uint8_t*)nullptr) : -1;
    e.in_off = op.a.p ? (long long)((uint8_t*)op.a.p - (uint8_t*)nullptr) : -1;
    e.stats_off = op.stats ? (long long)((uint8_t*)op.stats - (uint8_t*)nullptr) : -1;
    e.w_idx = op.w_idx;
    e.has_res = op.has_res ? 1 : 0;
  }
  return count;
}

extern "C" int crfr_resnet34_forward(int engine, const float* const* host_params, void* const* host_buffers,
                                     const crfr_resnet_io* io, void* ws, size_t ws_bytes, void* stream) {
  CRFR_TRY(check_io(io, "resnet34_forward"));
  CRFR_CHECK_ARG(host_params && ws, "resnet34_forward: null pointer");
  CRFR_CHECK_ARG(io->training || host_buffers, "resnet34_forward: eval mode needs the running statistics");
  Net net;
  init_net(net, engine, host_params, nullptr, host_buffers, io, ws, ws_bytes, (cudaStream_t)stream, true);
  net.forward();
  return net.err;
}

// ------------------------------------------------------------------------------------------------------------
// Residual knowledge-distillation step (distill_main.py:59-74, evaluated on one forward: SURVEY.md 8c-iii):
// teacher forward (eval), student + assistant forward (train), the six MSE terms on the NHWC bf16 features as they
// sit in the workspace (no fp32 NCHW round trip), and both backward passes - one native call.
// ------------------------------------------------------------------------------------------------------------
namespace {

__global__ void kd_total_kernel(const float* __restrict__ parts, float* __restrict__ losses) {
  losses[0] = parts[0];
  losses[1] = ((parts[1] + parts[2]) + (parts[3] + parts[4])) + parts[5];
}

struct KdLayout {
  size_t teacher, train, scratch, total;
};
KdLayout kd_layout(int batch, int size, int teacher_ir50 = 0) {
  KdLayout l;
  l.teacher = ((teacher_ir50 ? crfr_ir50_workspace_bytes(batch, size) : crfr_resnet34_workspace_bytes(batch, size, 0)) + 4095) &
              ~(size_t)4095;
  // + room for the second embedding-gradient slot of the student (L_s and L_a both reach s_emb)
  l.train = (crfr_resnet34_workspace_bytes(batch, size, 1) + (size_t)batch * kEmb * 2 + (4u << 20)) & ~(size_t)4095;
  l.scratch = 1 << 20;
  l.total = l.teacher + 2 * l.train + l.scratch;
  return l;
}

}  // namespace

extern "C" size_t crfr_kd_workspace_bytes(int batch, int size) {
  if (batch <= 0 || size != 112) return 0;
  return kd_layout(batch, size).total;
}
extern "C" size_t crfr_kd_workspace_bytes_ex(int batch, int size, int teacher_ir50) {
  if (batch <= 0 || size != 112) return 0;
  return kd_layout(batch, size, teacher_ir50).total;
}

extern "C" int crfr_kd_train_step(int engine, const float* const* teacher_params, void* const* teacher_buffers,
                                  const float* const* student_params, void* const* student_buffers,
                                  float* const* student_grads, const float* const* assistant_params,
                                  void* const* assistant_buffers, float* const* assistant_grads, const crfr_kd_io* kio,
                                  float* losses, void* ws, size_t ws_bytes, void* stream) {
  CRFR_CHECK_ARG(kio && kio->batch > 0 && kio->size == 112 && kio->x, "kd_train_step: batch/size/x invalid");
  CRFR_CHECK_ARG(teacher_params && teacher_buffers && student_params && student_grads && assistant_params &&
                     assistant_grads && losses && ws,
                 "kd_train_step: null pointer");
  const KdLayout lay = kd_layout(kio->batch, kio->size, kio->teacher_ir50);
  if (ws_bytes < lay.total) {
    crfr_set_error("kd_train_step: workspace %zu < %zu", ws_bytes, lay.total);
    return CRFR_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = (uint8_t*)ws;
  crfr_resnet_io io_eval = {}, io_train = {};
  io_eval.batch = io_train.batch = kio->batch;
  io_eval.size = io_train.size = kio->size;
  io_eval.x = kio->x;                                  // HR batch for the teacher
  io_train.x = kio->x_lr ? kio->x_lr : kio->x;         // LR batch (or the same one, as distill_main.py:59-61 feeds it)
  io_eval.momentum = io_train.momentum = kio->momentum;
  io_eval.eps = io_train.eps = kio->eps;
  io_eval.training = 0;
  io_train.training = 1;
  // The program of one step over three networks; real = false: dry run (arena offsets and weight-pack jobs only)
  auto program = [&](Net& T, Net& S, Net& A, bool real) -> int {
    if (kio->teacher_ir50) T.forward_ir50();
    else T.forward();
    if (!T.ok()) return T.err;
    S.forward();
    if (!S.ok()) return S.err;
    A.forward();
    if (!A.ok()) return A.err;

    float* parts = (float*)(base + lay.teacher + 2 * lay.train);   // [6] loss terms, then the reduction scratch
    void* lws = parts + 64;
    const size_t lws_bytes = lay.scratch - 64 * sizeof(float);
    const bool to_student = kio->assistant_grad_to_student != 0;
    auto grad_buf = [](Net& net, const Tensor& t) {
      bf16* g = (bf16*)net.alloc((size_t)t.n * t.h * t.w * t.ld * sizeof(bf16));
      net.add_slot(t, g, t.ld);
      return g;
    };
    // student loss: MSE(s_emb, t_emb.detach())
    {
      const long long numel = (long long)S.emb.n * S.emb.c;
      bf16* d_s = grad_buf(S, S.emb);
      if (real)
        CRFR_TRY(crfr_loss_kd(S.emb.p, nullptr, T.emb.p, numel, 0, 1.f, parts + 0, d_s, nullptr, nullptr, lws, lws_bytes, stream));
    }
    // assistant loss: sum_k MSE(t_k - s_k, a_k) over the four stage features and the embedding
    for (int k = 0; k < 5; ++k) {
      const Tensor& t = k < 4 ? T.feat[k] : T.emb;
      const Tensor& s_ = k < 4 ? S.feat[k] : S.emb;
      const Tensor& a = k < 4 ? A.feat[k] : A.emb;
      const long long numel = (long long)t.n * t.h * t.w * t.c;
      bf16* d_s = to_student ? grad_buf(S, s_) : nullptr;
      bf16* d_a = grad_buf(A, a);
      if (real)
        CRFR_TRY(crfr_loss_kd(t.p, s_.p, a.p, numel, 0, 1.f, parts + 1 + k, nullptr, d_s, d_a, lws, lws_bytes, stream));
    }
    if (!S.ok()) return S.err;
    if (!A.ok()) return A.err;
    if (real) {
      kd_total_kernel<<<1, 1, 0, st>>>(parts, losses);
      CRFR_COUNT_LAUNCH();
      CRFR_LAUNCH_CHECK();
    }
    S.backward();
    if (!S.ok()) return S.err;
    if (real && kio->events[0]) CRFR_CUDA(cudaEventRecord((cudaEvent_t)kio->events[0], st));   // student gradients are final
    A.backward();
    if (real && A.ok() && kio->events[1]) CRFR_CUDA(cudaEventRecord((cudaEvent_t)kio->events[1], st));
    return A.err;
  };
  {   // all weight packs of the step (three networks, both directions: ~180) as three launches
    std::vector<crfr_pack_job> jobs;
    Net T, S, A;
    init_net(T, engine, teacher_params, nullptr, teacher_buffers, &io_eval, base, lay.teacher, st, false);
    init_net(S, engine, student_params, student_grads, student_buffers, &io_train, base + lay.teacher, lay.train, st, false);
    init_net(A, engine, assistant_params, assistant_grads, assistant_buffers, &io_train, base + lay.teacher + lay.train,
             lay.train, st, false);
    T.collect = S.collect = A.collect = &jobs;
    CRFR_TRY(program(T, S, A, false));
    CRFR_TRY(crfr_pack_weight_batch(jobs.data(), (int)jobs.size(), st));
  }
  Net T, S, A;
  init_net(T, engine, teacher_params, nullptr, teacher_buffers, &io_eval, base, lay.teacher, st, true);
  init_net(S, engine, student_params, student_grads, student_buffers, &io_train, base + lay.teacher, lay.train, st, true);
  init_net(A, engine, assistant_params, assistant_grads, assistant_buffers, &io_train, base + lay.teacher + lay.train,
           lay.train, st, true);
  T.packs_done = S.packs_done = A.packs_done = true;
  return program(T, S, A, true);
}

extern "C" size_t crfr_ir50_workspace_bytes(int batch, int size) {
  if (batch <= 0 || size != 112) return 0;
  crfr_resnet_io io = {};
  io.batch = batch; io.size = size; io.training = 0; io.momentum = 0.1f; io.eps = 1e-5f;
  Net net;
  init_net(net, CRFR_ENGINE_AUTO, nullptr, nullptr, nullptr, &io, nullptr, ~(size_t)0 >> 2, nullptr, false);
  net.forward_ir50();
  return net.off + 65536;
}

extern "C" int crfr_ir50_forward(int engine, const float* const* host_params, void* const* host_buffers,
                                 const crfr_resnet_io* io, void* ws, size_t ws_bytes, void* stream) {
  CRFR_TRY(check_io(io, "ir50_forward"));
  CRFR_CHECK_ARG(host_params && host_buffers && ws, "ir50_forward: null pointer");
  CRFR_CHECK_ARG(!io->training, "ir50_forward: the teacher runs in eval mode only (train-mode Dropout is stochastic)");
  Net net;
  init_net(net, engine, host_params, nullptr, host_buffers, io, ws, ws_bytes, (cudaStream_t)stream, true);
  net.forward_ir50();
  return net.err;
}

extern "C" int crfr_resnet34_backward(int engine, const float* const* host_params, float* const* host_grads,
                                      const crfr_resnet_io* io, const float* d_emb, const float* const* d_feat,
                                      void* ws, size_t ws_bytes, void* stream) {
  CRFR_TRY(check_io(io, "resnet34_backward"));
  CRFR_CHECK_ARG(host_params && host_grads && ws, "resnet34_backward: null pointer");
  CRFR_CHECK_ARG(io->training, "resnet34_backward: the forward must have run in training mode");
  Net net;
  init_net(net, engine, host_params, host_grads, nullptr, io, ws, ws_bytes, (cudaStream_t)stream, false);
  net.forward();   // dry run: rebuild the tape and the saved-activation offsets
  if (!net.ok()) return net.err;
  net.exec = true;
  for (int l = 0; l < 4; ++l) net.seed_grad(net.feat[l], d_feat ? d_feat[l] : nullptr);
  net.seed_grad(net.emb, d_emb);
  if (!net.ok()) return net.err;
  net.backward();
  return net.err;
}
