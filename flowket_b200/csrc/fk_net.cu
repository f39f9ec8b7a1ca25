// Machine handle, layer-program builder, parameter transforms (weight normalisation, complex expansion),
// fp32 forward (log psi, conditional log probs) and backward (weighted / per-sample gradients).
//
// Reference semantics restated here (relative to /root/reference/src/flowket):
//   machines/conv_net_autoregressive_2D.py:24-74, machines/simple_conv_net_autoregressive_1D.py:8-64,
//   machines/complex_values_simple_conv_net_autoregressive_1D.py:11-59, deepar/layers/wrappers.py:123-134,
//   layers/complex/base_layer.py:18-35, layers/complex/tensorflow_ops.py:7-13.
#include <stdarg.h>

#include <algorithm>
#include <atomic>

#include "fk_net.cuh"

namespace fk {

static thread_local char g_error[1024] = "";

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// program builder
// ------------------------------------------------------------------------------------------------
struct Builder {
  fk_net* net;
  int64_t p_off = 0, e_off = 0;
  bool wn, complex_net;

  int new_buf(int channels) {
    net->bufs.push_back({channels, -1});
    return (int)net->bufs.size() - 1;
  }

  // adds a conv reading `in_buf`; taps given as (dh, dw) offsets
  ConvOp& add_conv(int in_buf, int cin, int out_buf, int out_coff, int cout, const std::vector<std::pair<int, int>>& taps,
                   int act, bool weight_norm) {
    ConvOp op;
    op.in_buf = in_buf; op.cin = cin; op.out_buf = out_buf; op.out_coff = out_coff; op.cout = cout;
    op.out2_buf = -1; op.res_buf = -1; op.pre_buf = -1; op.act = act;
    op.ntaps = (int)taps.size();
    for (int t = 0; t < op.ntaps; ++t) { op.dh[t] = taps[t].first; op.dw[t] = taps[t].second; }
    op.w_off = e_off; e_off += (int64_t)op.ntaps * cin * cout;
    op.b_off = e_off; e_off += cout;
    if (complex_net) {
      op.raw_cin = cin / 2; op.raw_cout = cout / 2;
      const int64_t ksz = (int64_t)op.ntaps * op.raw_cin * op.raw_cout;
      op.p_kernel = p_off; p_off += ksz;
      op.p_kernel_imag = p_off; p_off += ksz;
      op.p_bias = p_off; p_off += op.raw_cout;
      op.p_bias_imag = p_off; p_off += op.raw_cout;
      op.p_g = -1;
    } else {
      op.raw_cin = cin; op.raw_cout = cout;
      op.p_kernel = p_off; p_off += (int64_t)op.ntaps * cin * cout;
      op.p_bias = p_off; p_off += cout;
      op.p_g = -1;
      if (weight_norm) { op.p_g = p_off; p_off += cout; }
      op.p_kernel_imag = op.p_bias_imag = -1;
    }
    net->ops.push_back(op);
    return net->ops.back();
  }
};

static std::vector<std::pair<int, int>> taps_2d(int kh, int kw, int off_h, int off_w, int dil_w = 1) {
  std::vector<std::pair<int, int>> t;
  for (int a = 0; a < kh; ++a)
    for (int b = 0; b < kw; ++b) t.push_back({a - off_h, b * dil_w - off_w});
  return t;
}

static int build_conv2d(fk_net* net) {
  Builder B{net};
  B.wn = net->flags & FK_FLAG_WEIGHT_NORM;
  B.complex_net = false;
  const int C = net->C, k = net->k, pad = k - 1;
  const int nb = 2 * net->depth - 2;
  net->in_buf = B.new_buf(1);
  int v = net->in_buf, h = net->in_buf;
  int v_pair = -1, h_pair = -1;
  for (int b = 0; b < nb; ++b) {
    const bool last = (b == nb - 1);
    const bool res2 = (b >= 2 && b % 2 == 0 && !last);  // second block of a residual pair
    const int cin = (b == 0) ? 1 : C;
    if (b % 2 == 1 && !last) { v_pair = v; h_pair = h; }
    // vertical stack: k x k conv, ZeroPadding2D((pad,0),(pad/2,pad/2))
    const int a_buf = B.new_buf(C);  // relu(v')
    int v_next = a_buf;
    {
      if (res2) {
        v_next = B.new_buf(C);
        ConvOp& op = B.add_conv(v, cin, v_next, 0, C, taps_2d(k, k, pad, pad / 2), ACT_RELU, B.wn);
        op.res_buf = v_pair;
        op.out2_buf = a_buf;
      } else {
        B.add_conv(v, cin, a_buf, 0, C, taps_2d(k, k, pad, pad / 2), ACT_RELU, B.wn);
      }
    }
    // horizontal stack: 1 x k conv, ZeroPadding2D((0,0),(pad,0)), relu
    const int x1 = B.new_buf(C);
    B.add_conv(h, cin, x1, 0, C, taps_2d(1, k, 0, pad), ACT_RELU, B.wn);
    // 1x1 (C -> C/2) on x1 (RightShift first in the last block), relu -> concat[0:C/2]
    const int cc = B.new_buf(C);
    B.add_conv(x1, C, cc, 0, C / 2, {{0, last ? -1 : 0}}, ACT_RELU, B.wn);
    // 1x1 (C -> C/2) on DownShift(relu(v')), relu -> concat[C/2:C]
    B.add_conv(a_buf, C, cc, C / 2, C / 2, {{-1, 0}}, ACT_RELU, B.wn);
    // k x k conv on the concat, ZeroPadding2D((pad,0),(pad,0))
    const int hn = B.new_buf(C);
    {
      ConvOp& op = B.add_conv(cc, C, hn, 0, C, taps_2d(k, k, pad, pad), ACT_RELU, B.wn);
      if (res2) op.res_buf = h_pair;
    }
    v = v_next; h = hn;
  }
  net->logits_buf = B.new_buf(4);
  B.add_conv(h, C, net->logits_buf, 0, 4, {{0, 0}}, ACT_NONE, false);
  net->num_params = B.p_off; net->num_eff = B.e_off;
  return 0;
}

static std::vector<int> dilations(int n_layers, int max_dil) {
  std::vector<int> d;
  int cur = 1;
  for (int i = 0; i < n_layers; ++i) {
    d.push_back(cur);
    if (max_dil > 0 && cur < max_dil) cur *= 2;
  }
  return d;
}

static int build_conv1d(fk_net* net) {
  Builder B{net};
  B.wn = net->flags & FK_FLAG_WEIGHT_NORM;
  B.complex_net = false;
  const int C = net->C, k = net->k;
  net->in_buf = B.new_buf(1);
  int x = net->in_buf;
  auto dil = dilations(net->depth - 2, net->max_dil);
  for (int i = 0; i < net->depth - 2; ++i) {
    const int cin = (i == 0) ? 1 : C;
    const int out = B.new_buf(C);
    ConvOp& op = B.add_conv(x, cin, out, 0, C, taps_2d(1, k, 0, (k - 1) * dil[i], dil[i]), ACT_RELU, B.wn);
    if ((net->flags & FK_FLAG_SKIP) && i > 0) op.res_buf = x;
    x = out;
  }
  net->logits_buf = B.new_buf(4);
  B.add_conv(x, C, net->logits_buf, 0, 4, {{0, -1}}, ACT_NONE, B.wn);  // DownShift along the sequence
  net->num_params = B.p_off; net->num_eff = B.e_off;
  return 0;
}

static int build_cconv1d(fk_net* net) {
  Builder B{net};
  B.wn = false;
  B.complex_net = true;
  const int C = net->C, k = net->k;
  net->in_buf = B.new_buf(2);
  int x = net->in_buf;
  auto dil = dilations(net->depth - 1, net->max_dil);
  for (int i = 0; i < net->depth - 1; ++i) {
    const int cin = (i == 0) ? 2 : 2 * C;
    const int out = B.new_buf(2 * C);
    ConvOp& op = B.add_conv(x, cin, out, 0, 2 * C, taps_2d(1, k, 0, (k - 1) * dil[i], dil[i]), ACT_LNCOSH, false);
    op.pre_buf = B.new_buf(2 * C);
    x = out;
  }
  net->logits_buf = B.new_buf(4);
  B.add_conv(x, 2 * C, net->logits_buf, 0, 4, {{0, -1}}, ACT_NONE, false);
  net->num_params = B.p_off; net->num_eff = B.e_off;
  return 0;
}

static inline int64_t align4(int64_t x) { return (x + 3) & ~(int64_t)3; }

// liveness-based physical slot assignment for inference
static void assign_phys(fk_net* net) {
  const int nv = (int)net->bufs.size();
  std::vector<int> last_use(nv, -1), first_def(nv, 1 << 30);
  first_def[net->in_buf] = -1;
  for (int i = 0; i < (int)net->ops.size(); ++i) {
    const ConvOp& op = net->ops[i];
    last_use[op.in_buf] = i;
    if (op.res_buf >= 0) last_use[op.res_buf] = i;
    for (int b : {op.out_buf, op.out2_buf, op.pre_buf})
      if (b >= 0) { first_def[b] = std::min(first_def[b], i); last_use[b] = std::max(last_use[b], i); }
  }
  last_use[net->logits_buf] = 1 << 30;
  std::vector<int> slot_free_at;  // op index after which the slot is free
  int maxc = 1;
  for (int v = 0; v < nv; ++v) maxc = std::max(maxc, net->bufs[v].channels);
  // process buffers in order of definition
  std::vector<int> order(nv);
  for (int v = 0; v < nv; ++v) order[v] = v;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return first_def[a] < first_def[b]; });
  for (int v : order) {
    int chosen = -1;
    for (int s = 0; s < (int)slot_free_at.size(); ++s)
      if (slot_free_at[s] < first_def[v]) { chosen = s; break; }
    if (chosen < 0) { slot_free_at.push_back(0); chosen = (int)slot_free_at.size() - 1; }
    slot_free_at[chosen] = last_use[v];
    net->bufs[v].phys = chosen;
  }
  net->n_phys = (int)slot_free_at.size();
  // Slot and buffer sizes are carved in multiples of 4 floats so that every activation buffer starts 16-byte aligned
  // whatever the channel count, lattice size and chunk length are (float4 / cp.async accesses in the kernels).
  net->phys_channels = align4(maxc);
  int64_t tot = 0;
  for (int v = 0; v < nv; ++v) tot += align4(net->bufs[v].channels);
  net->train_floats_per_cfg = tot * net->sites;
}

int64_t infer_floats_per_cfg(const fk_net* net) { return (int64_t)net->n_phys * net->phys_channels * net->sites; }

void assign_infer_buffers(const fk_net* net, float* base, int64_t n, std::vector<float*>& ptrs) {
  ptrs.resize(net->bufs.size());
  const int64_t slot = n * net->sites * net->phys_channels;
  for (size_t v = 0; v < net->bufs.size(); ++v) ptrs[v] = base + slot * net->bufs[v].phys;
}

// ------------------------------------------------------------------------------------------------
// parameter transforms
// ------------------------------------------------------------------------------------------------
struct OpParam {
  long long p_kernel, p_bias, p_g, p_kernel_imag, p_bias_imag, w_off, b_off;
  int ntaps, cin, cout, raw_cin, raw_cout, mode;  // mode 0 plain, 1 WN linear g, 2 WN exp g, 3 complex
};

// wn_dir (optional, [num_eff], layout of weff): v / |v| of the weight-normalised kernels; wn_coef (optional, [op][2][64]):
// a[co] = s / |v| and gs[co] = d w / d g over v_hat -- what the fused Jacobian flush of fk_tc_grad.cu needs per output channel
__global__ void build_weff_kernel(const OpParam* __restrict__ table, const float* __restrict__ params,
                                  float* __restrict__ weff, float* __restrict__ weffT, float* __restrict__ wn_dir,
                                  float* __restrict__ wn_coef) {
  const OpParam o = table[blockIdx.x];
  __shared__ float scale[128];
  const int K = o.ntaps * o.cin;
  if (o.mode == 3) {
    // W = kr - i ki  ->  Wr = kr, Wi = -ki ; real block matrix over channels [re | im]
    const int rc = o.raw_cin, ro = o.raw_cout;
    for (int e = threadIdx.x; e < K * o.cout; e += blockDim.x) {
      const int co = e % o.cout, kci = e / o.cout;
      const int t = kci / o.cin, ci = kci % o.cin;
      const bool in_im = ci >= rc, out_im = co >= ro;
      const int rci = in_im ? ci - rc : ci, rco = out_im ? co - ro : co;
      const long long ridx = ((long long)t * rc + rci) * ro + rco;
      const float wr = params[o.p_kernel + ridx], wi = -params[o.p_kernel_imag + ridx];
      float v;
      if (!in_im && !out_im) v = wr;
      else if (in_im && !out_im) v = -wi;
      else if (!in_im && out_im) v = wi;
      else v = wr;
      weff[o.w_off + e] = v;
      weffT[o.w_off + ((long long)t * o.cout + co) * o.cin + ci] = v;
    }
    for (int co = threadIdx.x; co < o.cout; co += blockDim.x)
      weff[o.b_off + co] = co < ro ? params[o.p_bias + co] : -params[o.p_bias_imag + co - ro];
    return;
  }
  for (int co = threadIdx.x; co < o.cout; co += blockDim.x) {
    float sc = 1.f;
    if (o.mode != 0) {
      float sq = 0.f;
      for (int kk = 0; kk < K; ++kk) {
        const float v = params[o.p_kernel + (long long)kk * o.cout + co];
        sq = fmaf(v, v, sq);
      }
      const float g = params[o.p_g + co];
      const float inv = rsqrtf(fmaxf(sq, 1e-12f)), sg = (o.mode == 2 ? expf(g) : g);
      sc = inv * sg;
      if (wn_coef && co < 64) {
        wn_coef[(blockIdx.x * 2 + 0) * 64 + co] = sc;
        wn_coef[(blockIdx.x * 2 + 1) * 64 + co] = o.mode == 2 ? sg : 1.f;
      }
      if (wn_dir)
        for (int kk = 0; kk < K; ++kk) wn_dir[o.w_off + (long long)kk * o.cout + co] = params[o.p_kernel + (long long)kk * o.cout + co] * inv;
    }
    scale[co] = sc;
    weff[o.b_off + co] = params[o.p_bias + co];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * o.cout; e += blockDim.x) {
    const int co = e % o.cout, kci = e / o.cout;
    const int t = kci / o.cin, ci = kci % o.cin;
    const float v = params[o.p_kernel + e] * scale[co];
    weff[o.w_off + e] = v;
    weffT[o.w_off + ((long long)t * o.cout + co) * o.cin + ci] = v;
  }
}

// raw-parameter gradient from effective-weight gradient; grid (num_ops, nbatch)
__global__ void grad_transform_kernel(const OpParam* __restrict__ table, const float* __restrict__ params,
                                      const float* __restrict__ geff, long long geff_stride,
                                      float* __restrict__ graw, long long graw_stride) {
  const OpParam o = table[blockIdx.x];
  const float* ge = geff + (long long)blockIdx.y * geff_stride;
  float* gr = graw + (long long)blockIdx.y * graw_stride;
  const int K = o.ntaps * o.cin;
  if (o.mode == 0) {
    for (int e = threadIdx.x; e < K * o.cout; e += blockDim.x) gr[o.p_kernel + e] = ge[o.w_off + e];
    for (int co = threadIdx.x; co < o.cout; co += blockDim.x) gr[o.p_bias + co] = ge[o.b_off + co];
  } else if (o.mode == 3) {
    const int rc = o.raw_cin, ro = o.raw_cout;
    for (int e = threadIdx.x; e < o.ntaps * rc * ro; e += blockDim.x) {
      const int rco = e % ro, rest = e / ro;
      const int rci = rest % rc, t = rest / rc;
      auto G = [&](int ci, int co) { return ge[o.w_off + ((long long)t * o.cin + ci) * o.cout + co]; };
      const float dWr = G(rci, rco) + G(rci + rc, rco + ro);
      const float dWi = -G(rci + rc, rco) + G(rci, rco + ro);
      gr[o.p_kernel + e] = dWr;
      gr[o.p_kernel_imag + e] = -dWi;
    }
    for (int co = threadIdx.x; co < ro; co += blockDim.x) {
      gr[o.p_bias + co] = ge[o.b_off + co];
      gr[o.p_bias_imag + co] = -ge[o.b_off + ro + co];
    }
  } else {
    // W = v * inv * s,  inv = 1/|v|,  s = exp(g) | g
    for (int co = threadIdx.x; co < o.cout; co += blockDim.x) {
      float sq = 0.f, dot = 0.f;
      for (int kk = 0; kk < K; ++kk) {
        const float v = params[o.p_kernel + (long long)kk * o.cout + co];
        sq = fmaf(v, v, sq);
        dot = fmaf(v, ge[o.w_off + (long long)kk * o.cout + co], dot);
      }
      const float inv = rsqrtf(fmaxf(sq, 1e-12f));
      const float g = params[o.p_g + co];
      const float s = (o.mode == 2) ? expf(g) : g;
      const float vhat_dot = dot * inv;  // sum_k G v_hat
      gr[o.p_g + co] = vhat_dot * (o.mode == 2 ? s : 1.f);
      for (int kk = 0; kk < K; ++kk) {
        const long long idx = (long long)kk * o.cout + co;
        const float vh = params[o.p_kernel + idx] * inv;
        gr[o.p_kernel + idx] = s * inv * (ge[o.w_off + idx] - vh * vhat_dot);
      }
      gr[o.p_bias + co] = ge[o.b_off + co];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
static void fill_launch(const fk_net* net, const ConvOp& op, float* const* bp, int64_t n, ConvLaunch& a) {
  a.in = bp[op.in_buf]; a.in_cs = net->bufs[op.in_buf].channels; a.cin = op.cin;
  a.out = bp[op.out_buf]; a.out_cs = net->bufs[op.out_buf].channels; a.out_coff = op.out_coff; a.cout = op.cout;
  a.out2 = op.out2_buf >= 0 ? bp[op.out2_buf] : nullptr; a.out2_cs = op.cout;
  a.res = op.res_buf >= 0 ? bp[op.res_buf] : nullptr; a.res_cs = op.cout;
  a.pre = nullptr; a.pre_cs = op.cout;
  a.w = net->d_weff + op.w_off; a.bias = net->d_weff + op.b_off;
  a.ntaps = op.ntaps;
  for (int t = 0; t < op.ntaps; ++t) { a.dh[t] = op.dh[t]; a.dw[t] = op.dw[t]; }
  a.H = net->H; a.W = net->W; a.npos = n * net->sites;
  a.act = op.act; a.accumulate = 0;
}

int run_forward(fk_net* net, const int8_t* sigma, int64_t n, float* const* bp, cudaStream_t s) {
  if (launch_sigma_to_float(sigma, bp[net->in_buf], net->bufs[net->in_buf].channels, n * net->sites, s)) return 1;
  for (const ConvOp& op : net->ops) {
    ConvLaunch a;
    fill_launch(net, op, bp, n, a);
    if (op.act == ACT_LNCOSH) {
      // conv writes the pre-activation; lncosh applied by a separate elementwise pass
      float* pre = op.pre_buf >= 0 ? bp[op.pre_buf] : nullptr;
      FK_REQUIRE(pre != nullptr, "lncosh op without a pre-activation buffer");
      a.out = pre; a.out_cs = op.cout; a.out_coff = 0; a.act = ACT_NONE;
      if (launch_conv(a, s)) return 1;
      if (launch_lncosh(pre, bp[op.out_buf], op.cout, n * net->sites, s)) return 1;
    } else {
      if (launch_conv(a, s)) return 1;
    }
  }
  return 0;
}

// backward through the program.  act[v], grad[v]: per virtual buffer.  geff: effective-weight gradient
// (batch-summed: one vector; per-sample: n vectors with stride geff_stride).
static int run_backward(fk_net* net, int64_t n, float* const* act, float* const* grad, float* dz, float* geff,
                        int per_sample, int64_t geff_stride, cudaStream_t s) {
  const long long npos = n * net->sites;
  for (int i = (int)net->ops.size() - 1; i >= 0; --i) {
    const ConvOp& op = net->ops[i];
    const int out_cs = net->bufs[op.out_buf].channels;
    const float* pre = op.pre_buf >= 0 ? act[op.pre_buf] : nullptr;
    if (launch_dz(grad[op.out_buf], out_cs, op.out_coff, act[op.out_buf], out_cs, op.out_coff,
                  op.out2_buf >= 0 ? grad[op.out2_buf] : nullptr, op.out2_buf >= 0 ? act[op.out2_buf] : nullptr,
                  op.res_buf >= 0 ? grad[op.res_buf] : nullptr, pre, op.cout, op.act, dz, npos, s))
      return 1;
    if (launch_dw(act[op.in_buf], net->bufs[op.in_buf].channels, op.cin, dz, op.cout, op.ntaps, op.dh, op.dw,
                  net->H, net->W, n, geff + op.w_off, geff + op.b_off, per_sample, geff_stride, s))
      return 1;
    if (op.in_buf != net->in_buf) {
      ConvLaunch a;
      a.in = dz; a.in_cs = op.cout; a.cin = op.cout;
      a.out = grad[op.in_buf]; a.out_cs = net->bufs[op.in_buf].channels; a.out_coff = 0; a.cout = op.cin;
      a.out2 = nullptr; a.out2_cs = 0; a.res = nullptr; a.res_cs = 0; a.pre = nullptr; a.pre_cs = 0;
      a.w = net->d_weffT + op.w_off; a.bias = nullptr;
      a.ntaps = op.ntaps;
      for (int t = 0; t < op.ntaps; ++t) { a.dh[t] = -op.dh[t]; a.dw[t] = -op.dw[t]; }
      a.H = net->H; a.W = net->W; a.npos = npos; a.act = ACT_NONE; a.accumulate = 1;
      if (launch_conv(a, s)) return 1;
    }
  }
  return 0;
}

}  // namespace fk

using namespace fk;

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" const char* fk_last_error(void) { return fk::g_error; }
extern "C" int fk_version(void) { return 100; }
extern "C" int64_t fk_launch_count(void) { return fk::g_launches.load(); }

extern "C" int fk_net_create(fk_net_t** out, int kind, int H, int W, int depth, int channels, int kernel_size,
                             int max_dilation, int flags) {
  FK_REQUIRE(out != nullptr, "fk_net_create: out is NULL");
  FK_REQUIRE(kind >= 0 && kind <= 2, "fk_net_create: unknown machine kind %d", kind);
  FK_REQUIRE(H >= 1 && W >= 1 && channels >= 2 && channels % 2 == 0, "fk_net_create: bad shape H=%d W=%d C=%d", H, W, channels);
  FK_REQUIRE(kernel_size >= 1 && kernel_size % 2 == 1 && kernel_size * kernel_size <= MAX_TAPS,
             "fk_net_create: kernel_size %d not supported (odd, <= 3)", kernel_size);
  FK_REQUIRE(kind == FK_NET_CONV2D || H == 1, "fk_net_create: 1-D machines need H == 1");
  FK_REQUIRE((kind == FK_NET_CCONV1D ? 2 * channels : channels) <= 64, "fk_net_create: at most 64 real channels per layer");
  FK_REQUIRE(depth >= (kind == FK_NET_CONV1D ? 3 : 2), "fk_net_create: depth %d too small", depth);
  fk_net* net = new fk_net();
  net->kind = kind; net->H = H; net->W = W; net->depth = depth; net->C = channels; net->k = kernel_size;
  net->max_dil = max_dilation; net->flags = flags; net->sites = H * W;
  net->d_params = net->d_weff = net->d_weffT = nullptr; net->d_optable = nullptr; net->d_wn_dir = net->d_wn_coef = nullptr;
  net->d_tc_weights = nullptr; net->tc_weight_bytes = 0; net->d_tc_bwd = nullptr; net->d_tc_exact = nullptr; net->params_set = false;
  if (kind == FK_NET_CONV2D) build_conv2d(net);
  else if (kind == FK_NET_CONV1D) build_conv1d(net);
  else build_cconv1d(net);
  assign_phys(net);
  std::vector<OpParam> table;
  for (const ConvOp& op : net->ops) {
    OpParam o;
    o.p_kernel = op.p_kernel; o.p_bias = op.p_bias; o.p_g = op.p_g; o.p_kernel_imag = op.p_kernel_imag;
    o.p_bias_imag = op.p_bias_imag; o.w_off = op.w_off; o.b_off = op.b_off;
    o.ntaps = op.ntaps; o.cin = op.cin; o.cout = op.cout; o.raw_cin = op.raw_cin; o.raw_cout = op.raw_cout;
    o.mode = (kind == FK_NET_CCONV1D) ? 3 : (op.p_g < 0 ? 0 : ((flags & FK_FLAG_EXP_NORM) && kind == FK_NET_CONV2D ? 2 : 1));
    table.push_back(o);
  }
  cudaError_t e = cudaMalloc(&net->d_params, sizeof(float) * net->num_params);
  if (e == cudaSuccess) e = cudaMalloc(&net->d_weff, sizeof(float) * net->num_eff);
  if (e == cudaSuccess) e = cudaMalloc(&net->d_weffT, sizeof(float) * net->num_eff);
  if (e == cudaSuccess) e = cudaMalloc(&net->d_optable, sizeof(OpParam) * table.size());
  if (e == cudaSuccess) e = cudaMemcpy(net->d_optable, table.data(), sizeof(OpParam) * table.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(net->d_weffT, 0, sizeof(float) * net->num_eff);
  if (e == cudaSuccess && kind == FK_NET_CONV2D) {   // tables of the fused per-sample Jacobian flush (fk_tc_grad.cu)
    e = cudaMalloc(&net->d_wn_dir, sizeof(float) * net->num_eff);
    if (e == cudaSuccess) e = cudaMalloc(&net->d_wn_coef, sizeof(float) * 128 * net->ops.size());
    if (e == cudaSuccess) e = cudaMemset(net->d_wn_dir, 0, sizeof(float) * net->num_eff);
    if (e == cudaSuccess) e = cudaMemset(net->d_wn_coef, 0, sizeof(float) * 128 * net->ops.size());
  }
  if (e != cudaSuccess) {
    set_error("fk_net_create: CUDA allocation failed: %s", cudaGetErrorString(e));
    fk_net_destroy(net);
    return 1;
  }
  // every device allocation of the handle happens here (the tensor-core engines' weight images and wiring tables too)
  if (tc_prepare(net) || tc_grad_prepare(net) || tcx_prepare(net)) {
    fk_net_destroy(net);
    return 1;
  }
  *out = net;
  return 0;
}

extern "C" int fk_net_destroy(fk_net_t* net) {
  if (!net) return 0;
  cudaFree(net->d_params); cudaFree(net->d_weff); cudaFree(net->d_weffT); cudaFree(net->d_optable);
  cudaFree(net->d_tc_weights);
  cudaFree(net->d_tc_bwd);
  cudaFree(net->d_tc_exact);
  cudaFree(net->d_wn_dir); cudaFree(net->d_wn_coef);
  delete net;
  return 0;
}

extern "C" int fk_net_num_params(const fk_net_t* net, int64_t* num_params) {
  FK_REQUIRE(net && num_params, "fk_net_num_params: NULL argument");
  *num_params = net->num_params;
  return 0;
}

extern "C" int fk_net_set_params(fk_net_t* net, const float* params, void* stream) {
  FK_REQUIRE(net && params, "fk_net_set_params: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  FK_CHECK_CUDA(cudaMemcpyAsync(net->d_params, params, sizeof(float) * net->num_params, cudaMemcpyDeviceToDevice, s));
  build_weff_kernel<<<(unsigned)net->ops.size(), 128, 0, s>>>((const OpParam*)net->d_optable, net->d_params,
                                                               net->d_weff, net->d_weffT, net->d_wn_dir, net->d_wn_coef);
  FK_CHECK_LAUNCH();
  net->params_set = true;
  if (tc_supported(net)) {
    if (tc_pack_weights(net, s)) return 1;
    if (tc_grad_supported(net) && tc_grad_pack_weights(net, s)) return 1;
    if (tcx_supported(net) && tcx_pack_weights(net, s)) return 1;
  }
  return 0;
}

extern "C" int64_t fk_log_psi_workspace_bytes(const fk_net_t* net, int64_t n, int engine) {
  if (!net) return -1;
  if (engine == FK_ENGINE_TC || engine == FK_ENGINE_TC_EXACT) return tc_log_psi_workspace_bytes(net, n);
  return infer_floats_per_cfg(net) * (int64_t)sizeof(float) * std::max<int64_t>(n, 1);
}

static int forward_chunks(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, float* cond_out, void* ws,
                          int64_t ws_bytes, cudaStream_t s) {
  FK_REQUIRE(net->params_set, "machine parameters were never set (fk_net_set_params)");
  const int64_t per_cfg = infer_floats_per_cfg(net) * (int64_t)sizeof(float);
  int64_t chunk = ws_bytes / per_cfg;
  FK_REQUIRE(chunk >= 1, "workspace too small: %lld bytes, need at least %lld per configuration", (long long)ws_bytes,
             (long long)per_cfg);
  chunk = std::min<int64_t>(chunk, 1 << 20);
  std::vector<float*> bp;
  for (int64_t i = 0; i < n; i += chunk) {
    const int64_t m = std::min(chunk, n - i);
    assign_infer_buffers(net, (float*)ws, m, bp);
    if (run_forward(net, sigma + i * net->sites, m, bp.data(), s)) return 1;
    if (launch_head(bp[net->logits_buf], sigma + i * net->sites, net->sites, m, log_psi_out ? log_psi_out + 2 * i : nullptr,
                    cond_out ? cond_out + 2 * i * net->sites : nullptr, s))
      return 1;
  }
  return 0;
}

extern "C" int fk_log_psi(fk_net_t* net, const int8_t* sigma, int64_t n, float* log_psi_out, int engine, void* ws,
                          int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && (n == 0 || (sigma && log_psi_out && ws)), "fk_log_psi: NULL argument");
  if (n == 0) return 0;
  if (engine == FK_ENGINE_TC) {
    FK_REQUIRE(tc_supported(net), "fk_log_psi: the tensor-core engine supports ConvNetAutoregressive2D with 32 channels, kernel 3 only");
    return tc_log_psi(net, sigma, n, log_psi_out, ws, ws_bytes, (cudaStream_t)stream);
  }
  if (engine == FK_ENGINE_TC_EXACT) {
    FK_REQUIRE(tcx_supported(net), "fk_log_psi: the tc-exact engine supports ConvNetAutoregressive2D with 32 channels, kernel 3, lattices up to one 128-row tile");
    return tcx_log_psi(net, sigma, n, log_psi_out, (cudaStream_t)stream);
  }
  FK_REQUIRE(engine == FK_ENGINE_FP32, "fk_log_psi: unknown engine %d", engine);
  return forward_chunks(net, sigma, n, log_psi_out, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int fk_cond_log_probs(fk_net_t* net, const int8_t* sigma, int64_t n, float* out, void* ws, int64_t ws_bytes,
                                 void* stream) {
  FK_REQUIRE(net && (n == 0 || (sigma && out && ws)), "fk_cond_log_probs: NULL argument");
  if (n == 0) return 0;
  return forward_chunks(net, sigma, n, nullptr, out, ws, ws_bytes, (cudaStream_t)stream);
}

// ---- gradients ------------------------------------------------------------------------------------
// workspace layout per chunk of m configurations:
//   act  [train_floats_per_cfg * m] | grad [train_floats_per_cfg * m] | dz [sites * maxC * m] | coef [2 m]
//   | geff: batch-summed [num_eff]  or per-sample [m * num_eff]
static int64_t grad_bytes_per_cfg(const fk_net* net, int per_sample) {
  int64_t f = 2 * net->train_floats_per_cfg + (int64_t)net->sites * net->phys_channels + 2;
  if (per_sample) f += net->num_eff;
  return f * (int64_t)sizeof(float);
}

extern "C" int64_t fk_grad_workspace_bytes(const fk_net_t* net, int64_t B, int per_sample) {
  if (!net) return -1;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, per_sample ? 256 : 1024));
  return grad_bytes_per_cfg(net, per_sample) * chunk + (per_sample ? 0 : net->num_eff * (int64_t)sizeof(float)) + 256;
}

__global__ void weighted_coef_kernel(const float* __restrict__ y, long long n, float* __restrict__ cre,
                                     float* __restrict__ cim) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // L = sum_b 2 Re(log psi_b y_b) = sum_b (2 Re y_b) Re log psi_b + (-2 Im y_b) Im log psi_b
  cre[i] = 2.f * y[2 * i];
  cim[i] = -2.f * y[2 * i + 1];
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

static int grad_impl(fk_net* net, const int8_t* sigma, const float* y, int64_t B, float* grad_out, float* O_re,
                     float* O_im, void* ws, int64_t ws_bytes, cudaStream_t s) {
  FK_REQUIRE(net->params_set, "machine parameters were never set (fk_net_set_params)");
  const int per_sample = (y == nullptr);
  const int64_t per_cfg = grad_bytes_per_cfg(net, per_sample);
  FK_REQUIRE(reinterpret_cast<uintptr_t>(ws) % 16 == 0, "gradient workspace must be 16-byte aligned");
  const int64_t fixed = (per_sample ? 0 : net->num_eff * (int64_t)sizeof(float)) + 256;
  int64_t chunk = (ws_bytes - fixed) / per_cfg;
  FK_REQUIRE(chunk >= 1, "gradient workspace too small: %lld bytes (need %lld + %lld per configuration)",
             (long long)ws_bytes, (long long)fixed, (long long)per_cfg);
  chunk = std::min<int64_t>(chunk, 8192);  // dw_kernel grid limit: chunk*sites/4096 <= 65535
  const size_t nv = net->bufs.size();
  std::vector<float*> act(nv), grad(nv);
  float* base = (float*)ws;
  float* geff_sum = nullptr;
  if (!per_sample) {
    geff_sum = base;
    base += align4(net->num_eff);
    FK_CHECK_CUDA(cudaMemsetAsync(geff_sum, 0, sizeof(float) * net->num_eff, s));
  }
  for (int64_t i = 0; i < B; i += chunk) {
    const int64_t m = std::min(chunk, B - i);
    float* p = base;
    // every carve is a multiple of 4 floats (16-byte alignment of all buffers for any channel count / chunk)
    for (size_t v = 0; v < nv; ++v) { act[v] = p; p += align4(net->bufs[v].channels) * net->sites * m; }
    float* grad_base = p;
    for (size_t v = 0; v < nv; ++v) { grad[v] = p; p += align4(net->bufs[v].channels) * net->sites * m; }
    float* dz = p; p += (int64_t)net->sites * net->phys_channels * m;
    float* cre = p; p += align4(m);
    float* cim = p; p += align4(m);
    float* geff = per_sample ? p : geff_sum;
    const int8_t* sg = sigma + i * net->sites;
    if (run_forward(net, sg, m, act.data(), s)) return 1;
    const int passes = per_sample ? (O_im ? 2 : 1) : 1;
    for (int pass = 0; pass < passes; ++pass) {
      FK_CHECK_CUDA(cudaMemsetAsync(grad_base, 0, sizeof(float) * net->train_floats_per_cfg * m, s));
      if (per_sample) {
        fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(cre, m, pass == 0 ? 1.f : 0.f);
        fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(cim, m, pass == 0 ? 0.f : 1.f);
      } else {
        weighted_coef_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(y + 2 * i, m, cre, cim);
      }
      FK_CHECK_LAUNCH();
      if (launch_head_backward(act[net->logits_buf], sg, net->sites, m, cre, cim, grad[net->logits_buf], s)) return 1;
      if (run_backward(net, m, act.data(), grad.data(), dz, geff, per_sample, net->num_eff, s)) return 1;
      if (per_sample) {
        float* dst = (pass == 0 ? O_re : O_im) + i * net->num_params;
        grad_transform_kernel<<<dim3((unsigned)net->ops.size(), (unsigned)m), 128, 0, s>>>(
            (const OpParam*)net->d_optable, net->d_params, geff, net->num_eff, dst, net->num_params);
        FK_CHECK_LAUNCH();
      }
    }
  }
  if (!per_sample) {
    grad_transform_kernel<<<dim3((unsigned)net->ops.size(), 1), 128, 0, s>>>(
        (const OpParam*)net->d_optable, net->d_params, geff_sum, net->num_eff, grad_out, net->num_params);
    FK_CHECK_LAUNCH();
  }
  return 0;
}

namespace fk {
int64_t grad_transform_launch(fk_net* net, const float* geff, float* graw, cudaStream_t s) {
  grad_transform_kernel<<<dim3((unsigned)net->ops.size(), 1), 128, 0, s>>>((const OpParam*)net->d_optable, net->d_params, geff,
                                                                           net->num_eff, graw, net->num_params);
  FK_CHECK_LAUNCH();
  return 0;
}
int grad_transform_rows_launch(fk_net* net, const float* geff, float* graw, int64_t m, cudaStream_t s) {
  for (int64_t r0 = 0; r0 < m; r0 += 32768) {   // grid.y limit
    const int64_t rows = std::min<int64_t>(32768, m - r0);
    grad_transform_kernel<<<dim3((unsigned)net->ops.size(), (unsigned)rows), 128, 0, s>>>(
        (const OpParam*)net->d_optable, net->d_params, geff + r0 * net->num_eff, net->num_eff, graw + r0 * net->num_params,
        net->num_params);
    FK_CHECK_LAUNCH();
  }
  return 0;
}
}  // namespace fk

extern "C" int fk_grad_weighted(fk_net_t* net, const int8_t* sigma, const float* y, int64_t B, float* grad_out, void* ws,
                                int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma && y && grad_out && ws, "fk_grad_weighted: NULL argument");
  return grad_impl(net, sigma, y, B, grad_out, nullptr, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int fk_grad_per_sample(fk_net_t* net, const int8_t* sigma, int64_t B, float* O_re, float* O_im, void* ws,
                                  int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma && O_re && ws, "fk_grad_per_sample: NULL argument");
  return grad_impl(net, sigma, nullptr, B, nullptr, O_re, O_im, ws, ws_bytes, (cudaStream_t)stream);
}

// tensor-core engine of the per-sample Jacobians (fk_tc_grad.cu)
extern "C" int64_t fk_grad_per_sample_tc_workspace_bytes(const fk_net_t* net, int64_t B) {
  if (!net || !tc_grad_supported(net)) return -1;
  return tc_grad_per_sample_workspace_bytes(net, B);
}

extern "C" int fk_grad_per_sample_tc(fk_net_t* net, const int8_t* sigma, int64_t B, float* O_re, float* O_im, void* ws,
                                     int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma && O_re && ws, "fk_grad_per_sample_tc: NULL argument");
  FK_REQUIRE(tc_grad_supported(net), "fk_grad_per_sample_tc: supports ConvNetAutoregressive2D, 32 channels, kernel 3, lattices that fit one M tile");
  if (B == 0) return 0;
  return tc_grad_per_sample(net, sigma, B, O_re, O_im, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int64_t fk_jacobian_rows_tc_workspace_bytes(const fk_net_t* net, int64_t B) {
  if (!net || !tc_grad_supported(net)) return -1;
  return tc_jacobian_rows_workspace_bytes(net, B);
}

extern "C" int fk_jacobian_rows_tc(fk_net_t* net, const int8_t* sigma, int64_t B, void* X, int64_t rld, int64_t row_re,
                                   int64_t row_im, void* ws, int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma && X && ws, "fk_jacobian_rows_tc: NULL argument");
  FK_REQUIRE(tc_grad_supported(net), "fk_jacobian_rows_tc: supports ConvNetAutoregressive2D, 32 channels, kernel 3, lattices that fit one M tile");
  FK_REQUIRE(row_re >= 0 && row_re + B <= rld && (row_im < 0 || row_im + B <= rld), "fk_jacobian_rows_tc: rows out of range");
  if (B == 0) return 0;
  return tc_jacobian_rows(net, sigma, B, X, rld, row_re, row_im, ws, ws_bytes, (cudaStream_t)stream);
}

// tensor-core engine of the weighted gradient (fk_tc_grad.cu)
extern "C" int64_t fk_grad_weighted_tc_workspace_bytes(const fk_net_t* net, int64_t B) {
  if (!net || !tc_grad_supported(net)) return -1;
  return tc_grad_workspace_bytes(net, B);
}

extern "C" int fk_grad_weighted_tc(fk_net_t* net, const int8_t* sigma, const float* y, int64_t B, float* grad_out, void* ws,
                                   int64_t ws_bytes, void* stream) {
  FK_REQUIRE(net && sigma && y && grad_out && ws, "fk_grad_weighted_tc: NULL argument");
  FK_REQUIRE(tc_grad_supported(net), "fk_grad_weighted_tc: supports ConvNetAutoregressive2D, 32 channels, kernel 3, lattices that fit one M tile");
  if (B == 0) return 0;
  return tc_grad_weighted(net, sigma, y, B, grad_out, ws, ws_bytes, (cudaStream_t)stream);
}
