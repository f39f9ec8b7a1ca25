// Tensor-core gradient of  L = sum_b 2 Re(log psi(sigma_b) y_b)  for ConvNetAutoregressive2D (C = 32, k = 3, one M tile).
//
//   1. tc_forward_kernel (fk_tc.cu) in dump mode: fp16 activation tiles, relu masks and logits of every block;
//   2. tc_backward_kernel: backward-data through all blocks in ONE persistent kernel -- the gradient tiles are the A
//      operands (shifted views with negated tap offsets), the transposed kernels the B operands, relu masks / residual
//      fan-out / concat split in the epilogues; it writes the pre-activation gradients dz of every conv (fp16, scaled);
//   3. tc_dw_kernel: dW[(ci), (tap, co)] = sum_{cfg, p} x(p + off_tap)[ci] dz(p)[co] as tcgen05 MMAs with the lattice
//      positions as the K dimension (both operands MN-major views of the dumped tiles), accumulated over configurations
//      in TMEM, flushed once per work item with atomics; biases by a small reduction kernel;
//   4. grad_transform_kernel (fk_net.cu): effective-weight gradient -> raw (weight-normalised) parameter gradient.
//
// Numerics: fp16 operands, fp32 accumulation; the head gradient is scaled by a power of two (passed by the host) so
// that the backward tiles stay in the fp16 normal range, and the scale is removed in fp32 at the end.
// Semantics mirror run_backward() of fk_net.cu (the fp32 engine), which is the 1e-5 contract; tolerance of this
// engine: tests/test_gpu_tc.py.
#include <algorithm>
#include <vector>

#include <cuda_bf16.h>

#include "fk_net.cuh"
#include "fk_tc_common.cuh"

namespace fk {

// transposed weight image of a block (bytes)
constexpr int IMGB_HT = 0;          // 9 taps x 2 k-steps x 1024 B : B[n = concat ch][k = h' ch]
constexpr int IMGB_XXT = 18432;     // 1024 B                      : B[n = x1 ch (32)][k = 16]
constexpr int IMGB_YT = 19456;      // 1024 B                      : B[n = relu(v') ch (32)][k = 16]
constexpr int IMGB_XT = 20480;      // 3 x 2 x 1024 B              : B[n = h ch][k = x1 ch]
constexpr int IMGB_VT = 26624;      // 9 x 2 x 1024 B              : B[n = v ch][k = v' ch]
constexpr int IMGB_HEAD = 45056;    // fp32 [32][4] head kernel
constexpr int IMGB_BYTES = 45568;
constexpr int DZ_TILE = 8192;       // compact dz tile [4 cg][128 rows][8 ch] fp16
constexpr int NDUMP = 5;            // x1, relu(v'), residual v, concat, h_out

struct BwdPackDesc {
  long long w[5];      // V, X, XX, Y, H offsets into the effective-weight buffer
  long long w_head;
  int cin;
};

__global__ void tc_pack_bwd_kernel(const float* __restrict__ weff, const BwdPackDesc* __restrict__ pd,
                                   uint8_t* __restrict__ images) {
  const BwdPackDesc d = pd[blockIdx.x];
  uint8_t* img = images + (size_t)blockIdx.x * IMGB_BYTES;
  __half* w16 = reinterpret_cast<__half*>(img);
  // regions: (byte offset, taps, k-steps, forward op, forward cin, forward cout); B[n = ci][k = co]
  const int reg_off[5] = {IMGB_HT, IMGB_XXT, IMGB_YT, IMGB_XT, IMGB_VT};
  const int reg_taps[5] = {9, 1, 1, 3, 9};
  const int reg_ks[5] = {2, 1, 1, 2, 2};
  const int reg_op[5] = {4, 2, 3, 1, 0};
  const int reg_cout[5] = {32, 16, 16, 32, 32};
  for (int r = 0; r < 5; ++r) {
    const int cin = (reg_op[r] <= 1) ? d.cin : 32, cout = reg_cout[r];
    const int total = reg_taps[r] * reg_ks[r] * 2 * 32 * 8;   // (tap, kstep, kgroup, n = 32, e)
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int el = e & 7, n = (e >> 3) & 31, g = (e >> 8) & 1;
      const int ks = (e >> 9) % reg_ks[r], tap = (e >> 9) / reg_ks[r];
      const int co = ks * 16 + g * 8 + el;
      float v = 0.f;
      if (n < cin && co < cout) v = weff[d.w[reg_op[r]] + ((long long)tap * cin + n) * cout + co];
      w16[reg_off[r] / 2 + e] = __float2half_rn(v);
    }
  }
  float* hw = reinterpret_cast<float*>(img + IMGB_HEAD);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) hw[i] = weff[d.w_head + i];   // [ci][4]
}

// ---- TMEM store ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ================================================================================================================
// backward-data kernel
// ================================================================================================================
struct BwdArgs {
  const uint8_t* images;        // nb transposed weight images
  const uint32_t* mask;         // [cfg][nb][5][128]
  const float* logits;          // [cfg][128][4]
  const int8_t* sigma;          // [cfg][H*W]
  const float* coef;            // [cfg][2]: scaled (d L / d Re log psi, d L / d Im log psi)
  uint8_t* dz;                  // [cfg][nb][4][DZ_TILE]
  float* glog;                  // [cfg][128][4] scaled gradient w.r.t. the logits
  long long n;
  int H, W, P, nb, npos_g, p_first;
};

constexpr int BWD_NP = 2;

__global__ void __launch_bounds__(BWD_NP * 128 + 32, 1) tc_backward_kernel(BwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int pipe = tid >> 7, ltid = tid & 127;
  const bool is_producer = warp == BWD_NP * 4;
  const int tile_bytes = 64 * a.npos_g;
  uint8_t* wbuf = smem;
  uint8_t* g0 = smem + 2 * IMGB_BYTES;                                   // BWD_NP x 4 gradient tiles
  uint8_t* tail = g0 + (size_t)BWD_NP * 4 * tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);                    // full[0..1], empty[2..3], mma[4..5]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 64);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]);
  const uint32_t mbar = smem_u32(&bars[4 + (is_producer ? 0 : pipe)]);

  if (tid == 32) {
    mbar_init(full0, 1); mbar_init(full0 + 8, 1);
    mbar_init(empty0, BWD_NP); mbar_init(empty0 + 8, BWD_NP);
    for (int p = 0; p < BWD_NP; ++p) mbar_init(smem_u32(&bars[4 + p]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    uint4* z = reinterpret_cast<uint4*>(g0);
    const int n16 = BWD_NP * 4 * tile_bytes / 16;
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long groups = (a.n + BWD_NP - 1) / BWD_NP;
  const long long my_iters = (long long)blockIdx.x < groups ? (groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (is_producer) {
    const long long total = my_iters * a.nb;
    for (long long s = 0; s < total; ++s) {
      if ((tid & 31) == 0) {
        const uint32_t sel = (uint32_t)(s & 1);
        if (s >= 2) mbar_wait(empty0 + 8 * sel, (uint32_t)(((s >> 1) - 1) & 1));
        const int b = a.nb - 1 - (int)(s % a.nb);     // blocks in reverse
        mbar_expect_tx(full0 + 8 * sel, IMGB_BYTES);
        bulk_g2s(smem_u32(wbuf + (size_t)sel * IMGB_BYTES), a.images + (size_t)b * IMGB_BYTES, IMGB_BYTES, full0 + 8 * sel);
      }
      __syncwarp();
    }
  } else {
    const uint32_t tm_d = tmem_base + (uint32_t)(pipe * 128);        // MMA accumulators: 64 columns
    const uint32_t tm_ph = tm_d + 64, tm_pv = tm_d + 96;              // pending residual gradients (fp32)
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* gt = g0 + (size_t)pipe * 4 * tile_bytes;                 // tiles: 0 dzH, 1 dzc, 2 dzX, 3 dzV
    const uint32_t gt16 = smem_u32(gt) >> 4, tile16 = (uint32_t)(tile_bytes >> 4);
    const uint32_t kstep16 = 2u * (uint32_t)a.npos_g;
    const int HW = a.H * a.W;
    const uint32_t idesc32 = make_idesc(32);
    const uint64_t adesc0 = make_desc(0, a.npos_g, 8);
    const uint64_t bdesc32 = make_desc(0, 32, 8);
    const int bar_id = 1 + pipe;
    const int pos = a.p_first + ltid;
    const int r_ = pos / a.P - 2, c_ = pos % a.P - 2;
    const int site = (c_ >= 0 && r_ < a.H) ? r_ * a.W + c_ : -1;

    bool write_global = true;
    auto store_grad_row = [&](int tile, long long cfg, int b, const float* v) {
      uint8_t* sbase = gt + (size_t)tile * tile_bytes + (size_t)pos * 16;
      uint8_t* gbase = a.dz + (((size_t)cfg * a.nb + b) * 4 + tile) * DZ_TILE + (size_t)ltid * 16;
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        uint4 q;
        q.x = pack_h2(v[8 * cg + 0], v[8 * cg + 1]);
        q.y = pack_h2(v[8 * cg + 2], v[8 * cg + 3]);
        q.z = pack_h2(v[8 * cg + 4], v[8 * cg + 5]);
        q.w = pack_h2(v[8 * cg + 6], v[8 * cg + 7]);
        *reinterpret_cast<uint4*>(sbase + (size_t)cg * a.npos_g * 16) = q;
        if (write_global) *reinterpret_cast<uint4*>(gbase + (size_t)cg * 2048) = q;
      }
    };
    // D[col0..col0+32) (+)= sum over taps of view(tile, shift_t) x B_t   (K = 32: two k-steps; ksteps = 1: K = 16)
    auto issue = [&](int tile, int cg_off, const int* shifts, int ntaps, int ksteps, uint32_t w16, uint32_t col0) {
      const uint64_t ad0 = adesc0 + (uint64_t)(gt16 + (uint32_t)tile * tile16 + (uint32_t)cg_off * (uint32_t)a.npos_g + (uint32_t)a.p_first);
      uint32_t acc = 0;
      for (int t = 0; t < ntaps; ++t) {
        const uint64_t ad = ad0 + (uint64_t)(int64_t)shifts[t];
        const uint64_t bd = bdesc32 + w16 + (uint64_t)(t * ksteps) * 64;
        umma_f16(tm_d + col0, ad, bd, idesc32, acc);
        if (ksteps == 2) umma_f16(tm_d + col0, ad + kstep16, bd + 64, idesc32, 1u);
        acc = 1;
      }
    };

    uint32_t mma_phase = 0;
    long long step = 0;
    for (long long it = 0; it < my_iters; ++it) {
      const long long group = it * gridDim.x + blockIdx.x;
      const long long cfg = group * BWD_NP + pipe;
      const bool active = cfg < a.n;
      const long long cfgc = active ? cfg : 0;          // idle pipelines replay configuration 0 and write nothing
      write_global = active;
      float Gh[32], Gv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { Gh[i] = 0.f; Gv[i] = 0.f; }
      float g4[4] = {0.f, 0.f, 0.f, 0.f};
      if (site >= 0) {
        const float4 lg = *reinterpret_cast<const float4*>(a.logits + ((size_t)cfgc * 128 + ltid) * 4);
        const float x = 2.f * lg.x, y = 2.f * lg.y;
        const float m = fmaxf(x, y);
        const float ea = expf(x - m), eb = expf(y - m);
        const float p0 = ea / (ea + eb), p1 = eb / (ea + eb);
        const int sel = (1 - (int)a.sigma[cfgc * HW + site]) >> 1;
        const float cr = a.coef[2 * cfgc], ci = a.coef[2 * cfgc + 1];
        g4[0] = cr * ((sel == 0 ? 1.f : 0.f) - p0);
        g4[1] = cr * ((sel == 1 ? 1.f : 0.f) - p1);
        g4[2] = sel == 0 ? ci : 0.f;
        g4[3] = sel == 1 ? ci : 0.f;
      }
      if (active) *reinterpret_cast<float4*>(a.glog + ((size_t)cfg * 128 + ltid) * 4) = make_float4(g4[0], g4[1], g4[2], g4[3]);

      for (int kb = 0; kb < a.nb; ++kb, ++step) {
        const int b = a.nb - 1 - kb;
        const bool last = (b == a.nb - 1);
        const bool res2 = (b >= 2 && (b % 2) == 0 && !last);
        const bool fop = ((b % 2) == 1 && !last);          // first block of a residual pair
        const uint32_t wsel = (uint32_t)(step & 1);
        mbar_wait(full0 + 8 * wsel, (uint32_t)((step >> 1) & 1));
        const uint8_t* wimg = wbuf + (size_t)wsel * IMGB_BYTES;
        const uint32_t wimg16 = smem_u32(wimg) >> 4;
        const uint32_t* mk = a.mask + (((size_t)cfgc * a.nb + b) * NDUMP) * 128 + ltid;
        const uint32_t m_x1 = mk[0], m_a = mk[128], m_r = mk[256], m_c = mk[384], m_h = mk[512];
        if (last) {   // gradient w.r.t. the head input: W_head^T g
          const float* hw = reinterpret_cast<const float*>(wimg + IMGB_HEAD);
#pragma unroll
          for (int c = 0; c < 32; ++c)
            Gh[c] = hw[c * 4 + 0] * g4[0] + hw[c * 4 + 1] * g4[1] + hw[c * 4 + 2] * g4[2] + hw[c * 4 + 3] * g4[3];
        }
        // ---- phase 0: dz_H = G_h * [h_out > 0]; residual fan-out
        {
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = (m_h >> c) & 1u ? Gh[c] : 0.f;
          store_grad_row(0, cfgc, b, v);
          if (res2) tmem_st32(tm_ph + lane_sel, v);
        }
        fence_proxy_async();
        tc_fence_before();
        named_sync(bar_id, 128);
        // ---- phase 1: G_c = conv^T_H(dz_H)
        if (ltid < 32 && elect_one()) {   // (elected lane of a converged warp: descriptors stay in uniform registers)
          tc_fence_after();
          int sh[9];
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) sh[i * 3 + j] = (2 - i) * a.P + (2 - j);
          issue(0, 0, sh, 9, 2, wimg16 + IMGB_HT / 16, 0);
          umma_commit(mbar);
        }
        mbar_wait(mbar, mma_phase); mma_phase ^= 1;
        tc_fence_after();
        {
          float v[32];
          tmem_ld32(tm_d + lane_sel + 0, v);
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = (m_c >> c) & 1u ? v[c] : 0.f;
          store_grad_row(1, cfgc, b, v);
        }
        fence_proxy_async();
        tc_fence_before();
        named_sync(bar_id, 128);
        // ---- phase 2: G_x1 = dz_c[:, :16] W_XX^T (RightShift^T in the last block), G_a += shift(dz_c[:, 16:] W_Y^T)
        if (ltid < 32 && elect_one()) {   // (elected lane of a converged warp: descriptors stay in uniform registers)
          tc_fence_after();
          int s1[1] = {last ? 1 : 0};
          issue(1, 0, s1, 1, 1, wimg16 + IMGB_XXT / 16, 0);
          int s2[1] = {a.P};
          issue(1, 2, s2, 1, 1, wimg16 + IMGB_YT / 16, 32);
          umma_commit(mbar);
        }
        mbar_wait(mbar, mma_phase); mma_phase ^= 1;
        tc_fence_after();
        {
          float v[32];
          tmem_ld32(tm_d + lane_sel + 0, v);
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = (m_x1 >> c) & 1u ? v[c] : 0.f;
          store_grad_row(2, cfgc, b, v);
          tmem_ld32(tm_d + lane_sel + 32, v);
          if (res2) {
            float pv[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              pv[c] = (m_r >> c) & 1u ? Gv[c] : 0.f;
              v[c] = ((m_a >> c) & 1u ? v[c] : 0.f) + pv[c];
            }
            tmem_st32(tm_pv + lane_sel, pv);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = (m_a >> c) & 1u ? v[c] + Gv[c] : 0.f;
          }
          store_grad_row(3, cfgc, b, v);
        }
        fence_proxy_async();
        tc_fence_before();
        named_sync(bar_id, 128);
        // ---- phase 3: G_hin = conv^T_X(dz_X), G_vin = conv^T_V(dz_V)   (not needed below the first block)
        if (b > 0) {
          if (ltid < 32 && elect_one()) {   // (elected lane of a converged warp: descriptors stay in uniform registers)
            tc_fence_after();
            int sx[3] = {2, 1, 0};
            issue(2, 0, sx, 3, 2, wimg16 + IMGB_XT / 16, 0);
            int sv[9];
            for (int i = 0; i < 3; ++i)
              for (int j = 0; j < 3; ++j) sv[i * 3 + j] = (2 - i) * a.P + (1 - j);
            issue(3, 0, sv, 9, 2, wimg16 + IMGB_VT / 16, 32);
            umma_commit(mbar);
          }
          mbar_wait(mbar, mma_phase); mma_phase ^= 1;
          tc_fence_after();
          tmem_ld32(tm_d + lane_sel + 0, Gh);
          tmem_ld32(tm_d + lane_sel + 32, Gv);
          if (fop) {
            float pd[32];
            tmem_ld32(tm_ph + lane_sel, pd);
#pragma unroll
            for (int c = 0; c < 32; ++c) Gh[c] += pd[c];
            tmem_ld32(tm_pv + lane_sel, pd);
#pragma unroll
            for (int c = 0; c < 32; ++c) Gv[c] += pd[c];
          }
          tc_fence_before();
          named_sync(bar_id, 128);
        }
        if (ltid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * wsel) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ================================================================================================================
// weight-gradient kernel: positions are the K dimension
// ================================================================================================================
struct DwConv {
  int x_tile;       // dump tile index of the conv input
  int dz_tile;      // dz tile index (0..3) of the block
  int dz_cg;        // first channel group of dz (0, or 2 for the relu(v') half of the concat gradient)
  int ntaps, n;     // taps, output channels (32 or 16)
  int off[9];       // position offsets of the taps (forward direction), including the p_first base
  int col0;         // first TMEM column
  int cin;          // forward input channels (1 for the first block's V and X convs)
  long long w_off;  // offset of dW in the effective-weight gradient
  long long p_kernel, p_g;   // offsets of the raw kernel / weight-norm g in the parameter vector (p_g < 0: no weight norm)
  int op;                    // index of the convolution in the layer program (row of the weight-norm coefficient table)
};
struct DwUnit { int b, nconv; DwConv conv[3]; };
static_assert(sizeof(DwUnit) <= 320 && sizeof(DwUnit) % 8 == 0, "tc_dw_rows_kernel stages DwUnit in 320-byte shared-memory slots");

struct DwArgs2 {
  const uint8_t* dump; const uint8_t* dz; const DwUnit* units; float* geff;
  long long n; int nb, npos, num_units, cfg_chunk;
  // per-sample Jacobians: cfg_chunk == 1 and every configuration writes its own row of geff (row_stride floats apart,
  // plain stores, values multiplied by out_scale); row_stride == 0: one shared gradient, atomics
  long long row_stride;
  float out_scale;
  // Jacobian rows for the sample-space SR Gram (xrows != nullptr, cfg_chunk == 1): configuration cfg writes row
  // row_base + cfg of the panel-major bf16 matrix X[p / 64][rld rows][p % 64] in RAW parameters -- the weight-norm
  // transform (fk_net.cu::grad_transform_kernel) is applied in the flush, nothing passes through fp32 rows
  __nv_bfloat16* xrows;
  long long rld, row_base;
  const float* wn_dir;
  const float* wn_coef;
  int stages;      // operand ring depth (3; 2 in the row mode, whose v_hat cache takes the room)
};

// address (in elements) of parameter p of row r in the panel-major layout
__device__ __forceinline__ long long xrow_index(long long rld, long long r, long long p) {
  return ((p >> 6) * rld + r) * 64 + (p & 63);
}

constexpr int DW_VH_FLOATS = 13824;   // v_hat cache of the largest unit: V (9 taps) + X (3 taps), 32 x (32 + 4) floats per tap (tc_dw_rows_kernel)

// Flush of one tap's (ci, co) accumulator tile: TMEM lane = ci, N columns = co.  Compile-time N keeps the row in registers
// (a runtime trip count turns v[] into local memory -- measured: 1400 cycles per tap).
//   shared gradient: atomics;  per-sample row: the tile is contiguous in the row, so it is transposed through shared
//   memory and written with 512 contiguous bytes per store instruction.
template <int N>
__device__ __forceinline__ void dw_flush_tap(uint32_t taddr, float* dst, int cin, int lane, bool per_sample, float out_scale,
                                             float* stage) {
  float v[N];
  if (N == 32) tmem_ld32(taddr, reinterpret_cast<float(&)[32]>(v));
  else tmem_ld16(taddr, reinterpret_cast<float(&)[16]>(v));
  if (!per_sample) {
    if (lane < cin) {
#pragma unroll
      for (int co = 0; co < N; ++co) atomicAdd(dst + (long long)lane * N + co, v[co]);
    }
  } else if (cin == 32) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
      *reinterpret_cast<float4*>(stage + lane * 36 + 4 * q) =
          make_float4(v[4 * q] * out_scale, v[4 * q + 1] * out_scale, v[4 * q + 2] * out_scale, v[4 * q + 3] * out_scale);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const int e = i * 128 + lane * 4;
      *reinterpret_cast<float4*>(dst + e) = *reinterpret_cast<const float4*>(stage + (e / N) * 36 + (e % N));
    }
    __syncwarp();
  } else if (lane < cin) {   // first block: one input channel
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
      reinterpret_cast<float4*>(dst + (long long)lane * N)[q] =
          make_float4(v[4 * q] * out_scale, v[4 * q + 1] * out_scale, v[4 * q + 2] * out_scale, v[4 * q + 3] * out_scale);
  }
}

constexpr int DW_STAGES = 3;
// after the operand ring and the slack rows: barriers (128 B), 3 transpose buffers, cross-warp dots + coefficients, v_hat cache
constexpr int DW_TAIL_BYTES = 128 + 3 * 4608 + 4 * (192 + 192) + 256 + 8 * 320;   // ... + 8 unit-descriptor slots
constexpr int DW_ROWS_STAGES = 2;
#ifdef FK_DW_TRACE
__device__ long long fk_dw_trace_buf[64 * 8];
#define DWTRACE(i, k) do { if (blockIdx.x == 0 && (i) < 64 && (threadIdx.x & 31) == 0) fk_dw_trace_buf[(i) * 8 + (k)] = clock64(); } while (0)
#else
#define DWTRACE(i, k) do {} while (0)
#endif

__global__ void __launch_bounds__(384, 1) tc_dw_kernel(DwArgs2 a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xt_bytes = 64 * a.npos;
  const int stage_bytes = 3 * (xt_bytes + DZ_TILE);
  const int NST = a.stages;
  uint8_t* tail = smem + (size_t)NST * stage_bytes + 16 * a.npos * 16;   // slack: M rows 32..127 read past the tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);      // full[0..2], empty[3..5], done[6], tfree[7]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 64);
  // flush warps: 0 (always), 4 and 8 (row mode, 384 threads) -- all in TMEM lane quadrant 0 = the input channels
  const int nflush = blockDim.x >= 288 ? 3 : 1;
  const int fidx = warp >> 2;
  const bool is_flush = (warp & 3) == 0 && fidx < nflush;
  float* flush_stage = reinterpret_cast<float*>(tail + 128) + fidx * 1152;   // 32 x 36 floats per flush warp: transpose buffer
  // unit descriptors of the items in flight, in shared memory (slot = item counter & 7): a by-value copy per thread would live on
  // the local-memory stack (conv[k] / off[t] are indexed at run time), whose L1 is all but gone at this shared-memory carve-out
  DwUnit* sunit = reinterpret_cast<DwUnit*>(tail + DW_TAIL_BYTES - 8 * 320);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[DW_STAGES]), done = smem_u32(&bars[2 * DW_STAGES]);
  if (tid == 32) {
    for (int i = 0; i < DW_STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(done, 1);
    mbar_init(smem_u32(&bars[2 * DW_STAGES + 1]), nflush);   // tfree
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // the slack region is read as garbage M rows: keep it finite
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (NST * stage_bytes + 16 * a.npos * 16) / 16;
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const size_t dump_cfg_bytes = (size_t)(a.nb * NDUMP + 1) * xt_bytes;
  // A: x tile, MN-major: 8 channels contiguous (16 B), channel groups npos*16 B apart, positions 16 B apart,
  //    groups of 8 positions 128 B apart.  B: dz tile, MN-major, channel groups 2048 B apart.
  const uint64_t adesc0 = make_desc(0, 8, a.npos);
  const uint64_t bdesc0 = make_desc(0, 8, 128);

  const int chunks = (int)((a.n + a.cfg_chunk - 1) / a.cfg_chunk);
  const int items = a.num_units * chunks;
  // Three roles, each walking the same item list on its own: warp 2 loads (runs ahead through the stage ring), warp 1
  // issues the MMAs of an item once the previous item's accumulators have been flushed (`tfree`), warp 0 (TMEM lanes
  // 0..31 = input channels) flushes an item when its MMAs have retired (`done`).
  const uint32_t tfree = smem_u32(&bars[2 * DW_STAGES + 1]);
  if (warp == 2) {
    // ---- producer: one stage per configuration
    {
      uint32_t empty_phase = 0;   // bit i = parity of stage i
      long long prod_count = 0, pitem = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++pitem) {
        const int ui = item % a.num_units, ch = item / a.num_units;
        {   // the item's unit descriptor -> its shared-memory slot (whole warp); the other roles see it through `full`
          const uint32_t* src = reinterpret_cast<const uint32_t*>(a.units + ui);
          uint32_t* dst = reinterpret_cast<uint32_t*>(sunit + (pitem & 7));
          for (int w = lane; w < (int)(sizeof(DwUnit) / 4); w += 32) dst[w] = __ldg(src + w);
          __syncwarp();
        }
        if (lane != 0) continue;
        const DwUnit& u = sunit[pitem & 7];
        const long long c_beg = (long long)ch * a.cfg_chunk;
        const long long c_end = c_beg + a.cfg_chunk < a.n ? c_beg + a.cfg_chunk : a.n;
        for (long long cfg = c_beg; cfg < c_end; ++cfg) {
          const int st = (int)(prod_count % NST);
          if (prod_count >= NST) { mbar_wait(empty0 + 8 * st, (empty_phase >> st) & 1u); empty_phase ^= 1u << st; }
          uint8_t* sb = smem + (size_t)st * stage_bytes;
          mbar_expect_tx(full0 + 8 * st, (uint32_t)u.nconv * (xt_bytes + DZ_TILE));
          for (int k = 0; k < u.nconv; ++k) {
            bulk_g2s(smem_u32(sb + (size_t)k * xt_bytes), a.dump + (size_t)cfg * dump_cfg_bytes + (size_t)u.conv[k].x_tile * xt_bytes,
                     xt_bytes, full0 + 8 * st);
            bulk_g2s(smem_u32(sb + (size_t)3 * xt_bytes + (size_t)k * DZ_TILE),
                     a.dz + (((size_t)cfg * a.nb + u.b) * 4 + u.conv[k].dz_tile) * DZ_TILE, DZ_TILE, full0 + 8 * st);
          }
          ++prod_count;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer (elected lane of the converged warp: the descriptors stay in uniform registers)
    uint32_t full_phase = 0;
    long long cons_count = 0, item_count = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_count) {
      const int ch = item / a.num_units;
      const DwUnit& u = sunit[item_count & 7];   // (valid once the item's first `full` barrier has been seen)
      const long long c_beg = (long long)ch * a.cfg_chunk;
      const long long c_end = c_beg + a.cfg_chunk < a.n ? c_beg + a.cfg_chunk : a.n;
      DWTRACE(item_count, 0);
      mbar_wait(tfree, (uint32_t)((item_count & 1) ^ 1));   // previous item flushed (passes at once for the first item)
      tc_fence_after();
      DWTRACE(item_count, 1);
      for (long long cfg = c_beg; cfg < c_end; ++cfg) {
        const int st = (int)(cons_count % NST);
        mbar_wait(full0 + 8 * st, (full_phase >> st) & 1u); full_phase ^= 1u << st;
        tc_fence_after();
        DWTRACE(item_count, 2);
        if (elect_one()) {
          const uint32_t sb16 = smem_u32(smem + (size_t)st * stage_bytes) >> 4;
          const uint32_t first = cfg > c_beg ? 1u : 0u;
          for (int k = 0; k < u.nconv; ++k) {
            // (everything the MMA loop needs in registers: the asm memory clobber would reload it from the local copy)
            const int ntaps = u.conv[k].ntaps, n = u.conv[k].n, col0 = u.conv[k].col0;
            const uint32_t idesc = make_idesc(n) | (1u << 15) | (1u << 16);   // A and B MN-major
            const uint32_t x16 = sb16 + (uint32_t)k * (uint32_t)(xt_bytes >> 4);
            const uint64_t bd0 = bdesc0 + (uint64_t)(sb16 + (uint32_t)(3 * (xt_bytes >> 4)) + (uint32_t)k * (DZ_TILE >> 4) +
                                                     (uint32_t)u.conv[k].dz_cg * 128u);
            for (int t = 0; t < ntaps; ++t) {
              const uint32_t dcol = tmem + (uint32_t)(col0 + t * n);
              const uint64_t ad0 = adesc0 + (uint64_t)(x16 + (uint32_t)u.conv[k].off[t]);
              umma_f16(dcol, ad0, bd0, idesc, first);
#pragma unroll
              for (int ks = 1; ks < 8; ++ks) umma_f16(dcol, ad0 + (uint64_t)(16 * ks), bd0 + (uint64_t)(16 * ks), idesc, 1u);
            }
          }
          umma_commit(empty0 + 8 * st);                 // stage free when its MMAs retire
          if (cfg + 1 == c_end) umma_commit(done);
        }
        __syncwarp();
        DWTRACE(item_count, 3);
        ++cons_count;
      }
    }
  } else if (is_flush) {
    // ---- flush: the flush warps own TMEM lanes 0..31 = input channels
    uint32_t done_phase = 0;
    long long fitem = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++fitem) {
      const int ch = item / a.num_units;
      const DwUnit& u = sunit[fitem & 7];
      const long long c_beg = (long long)ch * a.cfg_chunk;
      mbar_wait(done, done_phase); done_phase ^= 1;
      tc_fence_after();
      DWTRACE(fitem, 4);
      for (int k = 0; k < u.nconv; ++k) {
        const int ntaps = u.conv[k].ntaps, n = u.conv[k].n, col0 = u.conv[k].col0, cin = u.conv[k].cin;
        const long long w_off = u.conv[k].w_off;
        for (int t = 0; t < ntaps; ++t) {
          float* dst = a.geff + c_beg * a.row_stride + w_off + (long long)t * cin * n;   // (c_beg * 0 in the shared mode)
          if (n == 32) dw_flush_tap<32>(tmem + (uint32_t)(col0 + t * 32), dst, cin, lane, a.row_stride != 0, a.out_scale, flush_stage);
          else dw_flush_tap<16>(tmem + (uint32_t)(col0 + t * 16), dst, cin, lane, a.row_stride != 0, a.out_scale, flush_stage);
        }
      }
      tc_fence_before();
      __syncwarp();
      DWTRACE(fitem, 5);
      if (lane == 0) mbar_arrive(tfree);   // the accumulators may be overwritten
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ================================================================================================================
// Jacobian-row variant of the weight-gradient kernel (fk_jacobian_rows_tc): one configuration per item, its accumulators
// leave TMEM as one bf16 row of X.  In tc_dw_kernel the MMAs of item i + 1 cannot start before item i has been flushed
// (one unit fills up to 384 of the 512 TMEM columns, so there is no second accumulator set), and the flush -- two passes
// over TMEM for the weight-norm transform, by the three warps that own lane quadrant 0 -- took as long as the MMAs
// (measured with tools/dw_rows_trace.py: 8.9 k + 9.7 k cycles per item).  Here the accumulators are first DRAINED to a
// shared-memory staging buffer (three warps), which frees TMEM for the next item at once; seven worker warps then turn
// the staged tile into the row -- any warp can read shared memory, TMEM lanes 0..31 only quadrant 0 -- while the tensor
// core already works on the next item.  Measured: 125.0 -> 115.8 ms per 8192 samples (Re + Im rows).  The overlap is
// worth less than the trace suggested because both halves live on the shared-memory data pipe: the MN-major operand
// fetch of the M = 128 MMAs (of which only the 32 input-channel rows are real) keeps it busy ~8 k cycles per item and
// starves the workers' loads (their two passes stretch from ~1.5 k to ~14 k cycles, tools/dw_rows_trace.py).  What would
// help is less operand traffic per configuration -- e.g. four configurations stacked along M and N (block-diagonal
// product, 2 KB instead of 5 KB per configuration and k-step) -- which needs a different TMEM budget; not built.
//   roles (512 threads): warp 2 producer | MMA issuers: warps 1, 12 | drainers: warps 0, 4 (TMEM quadrant 0) and 5, 9 (quadrant 1) |
//                        workers: warps 3, 6, 7, 8, 10, 11, 13, 14, 15
//   barriers: full/empty[stage] (operand ring), done (MMAs of the item retired), tfree (TMEM drained, 3 arrivals),
//             sfull (staging written, 3 arrivals), sfree (staging consumed, 7 arrivals)
// ================================================================================================================
// M = 64 MMAs: the A tile is the input-channel axis, of which 32 rows are real -- M = 64 halves the operand bytes the
// tensor core fetches for it (8 channel groups instead of 16).  Accumulator layout of a cta_group::1 M = 64 MMA: row r lives
// in TMEM lane 32 (r / 16) + r % 16, so the 32 real rows sit in lanes 0..15 of quadrants 0 and 1 -- two drainers per quadrant.
// Measured: 116.0 ms per 8192 samples with either M (the MN-major MMAs retire at ~75 cycles each regardless of the A bytes), so
// this only takes load off the shared-memory pipe; both variants pass the parity tests.
constexpr bool DWR_M64 = true;
constexpr int DWR_DRAINERS = DWR_M64 ? 4 : 3;
constexpr int DWR_WORKERS = DWR_M64 ? 9 : 7;
constexpr int DWR_ISSUERS = 2;                           // warps 1, 12 (.. 11 + DWR_ISSUERS - 1)
constexpr int DWR_THREADS = 512;
constexpr int DWR_SLOTS = 12;                            // taps of the largest unit (V: 9 + X: 3)
constexpr int DWR_STAGING_FLOATS = DWR_SLOTS * 32 * 36;  // [tap slot][ci][36]: rows padded for conflict-free 128-bit access
// after the operand ring: barriers (128 B) | staging | v_hat cache | coefficients [3][64] | partial dots [2][3][7][32] |
// dot totals [2][3][32]
constexpr int DWR_TAIL_BYTES = 128 + 4 * (DWR_STAGING_FLOATS + DW_VH_FLOATS + 192 + 2 * 3 * DWR_WORKERS * 32 + 2 * 3 * 32) + 4 * 320;   // + 4 unit descriptors

__global__ void __launch_bounds__(DWR_THREADS, 1) tc_dw_rows_kernel(DwArgs2 a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xt_bytes = 64 * a.npos;
  const int stage_bytes = 3 * (xt_bytes + DZ_TILE);
  const int NST = DW_ROWS_STAGES;
  // no slack region: the M rows 32..127 of an A tile (channel groups 4..15) read whatever follows the tile -- the other
  // stage, the staging buffer, the v_hat cache; they only reach TMEM lanes 32..127, which nobody reads
  uint8_t* tail = smem + (size_t)NST * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);      // full[0..1], empty[2..3], done[4], tfree[5], sfull[6], sfree[7]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 64);
  float* stg = reinterpret_cast<float*>(tail + 128);
  float* vh_s = stg + DWR_STAGING_FLOATS;
  float* cf_s = vh_s + DW_VH_FLOATS;                        // [3 convs][a[32] | gs[32]]
  float* xdot = cf_s + 192;                                 // [2 parities][3 convs][workers][32]
  float* wdot = xdot + 2 * 3 * DWR_WORKERS * 32;            // [2 parities][3 convs][32]: the dot totals of the item
  // Unit descriptors in shared memory, one slot per item in flight (worker on item i - 1 ... producer on item i + 2).  A by-value
  // copy per thread (as in tc_dw_kernel) lives on the local-memory stack because conv[k] / off[t] are indexed at run time, and with
  // 223 KB of the SM's 256 KB carved out as shared memory the L1 that backs the stack is almost gone: every field read in the
  // workers' inner loops went to L2 (measured: 14 k cycles per item in the two worker passes, tools/dw_rows_trace.py).
  DwUnit* sunit = reinterpret_cast<DwUnit*>(wdot + 2 * 3 * 32);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]), done = smem_u32(&bars[4]), tfree = smem_u32(&bars[5]),
                 sfull = smem_u32(&bars[6]), sfree = smem_u32(&bars[7]);
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, DWR_ISSUERS); }
    mbar_init(done, DWR_ISSUERS);   // every issuer commits its own MMAs
    mbar_init(tfree, DWR_DRAINERS);
    mbar_init(sfull, DWR_DRAINERS);
    mbar_init(sfree, DWR_WORKERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // everything an A tile can alias must be finite-or-harmless and initialised: zero the ring, the staging buffer and the cache
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = NST * stage_bytes / 16;
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    uint4* z2 = reinterpret_cast<uint4*>(stg);
    const int m16 = (DWR_STAGING_FLOATS + DW_VH_FLOATS) / 4;
    for (int i = tid; i < m16; i += blockDim.x) z2[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const size_t dump_cfg_bytes = (size_t)(a.nb * NDUMP + 1) * xt_bytes;
  const uint64_t adesc0 = make_desc(0, 8, a.npos);
  const uint64_t bdesc0 = make_desc(0, 8, 128);
  const long long items = (long long)a.num_units * a.n;    // unit-major: the CTAs in flight write adjacent rows of the same panels

  // drainer (quadrant dq, index dd among the drainers of the quadrant) / worker index
  bool is_drain;
  int dq = 0, dd = 0, dn = 1, widx = -1;
  if (DWR_M64) {
    is_drain = warp == 0 || warp == 4 || warp == 5 || warp == 9;
    dq = warp & 3; dd = (warp == 0 || warp == 5) ? 0 : 1; dn = 2;
    const int wmap[16] = {-1, -1, -1, 0, -1, -1, 1, 2, 3, -1, 4, 5, -1, 6, 7, 8};
    widx = wmap[warp];
  } else {
    is_drain = (warp & 3) == 0;
    dd = warp >> 2; dn = 3;
    if (warp == 3) widx = 0; else if (warp >= 5 && warp <= 7) widx = warp - 4; else if (warp >= 9 && warp <= 11) widx = warp - 5;
  }

  if (warp == 2) {
    // ---- producer: one stage per item
    uint32_t empty_phase = 0;
    long long count = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++count) {
      const long long cfg = item % a.n;
      const int st = (int)(count % NST);
      if (count >= NST) { mbar_wait(empty0 + 8 * st, (empty_phase >> st) & 1u); empty_phase ^= 1u << st; }
      // the item's unit descriptor -> its shared-memory slot (whole warp), visible to the other roles through the `full` barrier
      {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.units + item / a.n);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sunit + (count & 3));
        for (int w = lane; w < (int)(sizeof(DwUnit) / 4); w += 32) dst[w] = __ldg(src + w);
        __syncwarp();
      }
      if (lane == 0) {
        const DwUnit& u = sunit[count & 3];
        uint8_t* sb = smem + (size_t)st * stage_bytes;
        mbar_expect_tx(full0 + 8 * st, (uint32_t)u.nconv * (xt_bytes + DZ_TILE));
        for (int k = 0; k < u.nconv; ++k) {
          bulk_g2s(smem_u32(sb + (size_t)k * xt_bytes), a.dump + (size_t)cfg * dump_cfg_bytes + (size_t)u.conv[k].x_tile * xt_bytes,
                   xt_bytes, full0 + 8 * st);
          bulk_g2s(smem_u32(sb + (size_t)3 * xt_bytes + (size_t)k * DZ_TILE),
                   a.dz + (((size_t)cfg * a.nb + u.b) * 4 + u.conv[k].dz_tile) * DZ_TILE, DZ_TILE, full0 + 8 * st);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 || (warp >= 12 && warp < 11 + DWR_ISSUERS)) {
    // ---- MMA issuers (elected lane of a converged warp each).  One thread issues one tcgen05.mma per ~82 cycles whatever the
    // operands (tools/umma_bench3.cu: 82 / 41 / 24 cycles per M = 64, N = 32 MMA with 1 / 2 / 4 issuing warps -- the last is
    // the 128 B/clk operand stream), so the taps of an item are dealt to DWR_ISSUERS warps; a tap's eight k-steps stay with
    // one issuer (same accumulator columns, program order), every issuer commits its own MMAs to `empty` and `done`.
    const int iss = __shfl_sync(0xffffffffu, warp == 1 ? 0 : warp - 11, 0);
    uint32_t full_phase = 0;
    long long count = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++count) {
      const DwUnit& u = sunit[count & 3];   // (written by the producer before the item's TMA: valid once `full` has been seen)
      if (iss == 0) DWTRACE(count, 0);
      mbar_wait(tfree, (uint32_t)((count & 1) ^ 1));   // accumulators of the previous item drained (passes at once for the first)
      tc_fence_after();
      if (iss == 0) DWTRACE(count, 1);
      const int st = (int)(count % NST);
      mbar_wait(full0 + 8 * st, (full_phase >> st) & 1u); full_phase ^= 1u << st;
      tc_fence_after();
      if (iss == 0) DWTRACE(count, 2);
      if (elect_one()) {
        const uint32_t sb16 = smem_u32(smem + (size_t)st * stage_bytes) >> 4;
        int tap_index = 0;
        for (int k = 0; k < u.nconv; ++k) {
          const int ntaps = u.conv[k].ntaps, n = u.conv[k].n, col0 = u.conv[k].col0;
          uint32_t idesc = make_idesc(n) | (1u << 15) | (1u << 16);   // A and B MN-major
          if (DWR_M64) idesc = (idesc & ~(0x1Fu << 24)) | ((64u >> 4) << 24);
          const uint32_t x16 = sb16 + (uint32_t)k * (uint32_t)(xt_bytes >> 4);
          const uint64_t bd0 = bdesc0 + (uint64_t)(sb16 + (uint32_t)(3 * (xt_bytes >> 4)) + (uint32_t)k * (DZ_TILE >> 4) +
                                                   (uint32_t)u.conv[k].dz_cg * 128u);
          for (int t = 0; t < ntaps; ++t, ++tap_index) {
            if (tap_index % DWR_ISSUERS != iss) continue;
            const uint32_t dcol = tmem + (uint32_t)(col0 + t * n);
            const uint64_t ad0 = adesc0 + (uint64_t)(x16 + (uint32_t)u.conv[k].off[t]);
            umma_f16(dcol, ad0, bd0, idesc, 0u);
#pragma unroll
            for (int ks = 1; ks < 8; ++ks) umma_f16(dcol, ad0 + (uint64_t)(16 * ks), bd0 + (uint64_t)(16 * ks), idesc, 1u);
          }
        }
        umma_commit(empty0 + 8 * st);
        umma_commit(done);
      }
      __syncwarp();
      if (iss == 0) DWTRACE(count, 3);
    }
  } else if (is_drain) {
    // ---- drainers (TMEM lane quadrant 0 = the input channels): accumulators -> staging, then TMEM is free
    long long count = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++count) {
      const DwUnit& u = sunit[count & 3];
      mbar_wait(done, (uint32_t)(count & 1));
      tc_fence_after();
      mbar_wait(sfree, (uint32_t)((count & 1) ^ 1));   // the workers are through with the previous item's staging
#ifndef FK_DW_TRACE_WORKER
      DWTRACE(count, 4);
#endif
      int slot = 0;
      for (int k = 0; k < u.nconv; ++k) {
        const int ntaps = u.conv[k].ntaps, n = u.conv[k].n, col0 = u.conv[k].col0;
        for (int t = 0; t < ntaps; ++t, ++slot) {
          if (slot % dn != dd) continue;
          const uint32_t tq = tmem + ((uint32_t)(dq * 32) << 16);          // this warp's TMEM lane quadrant
          const int row = DWR_M64 ? dq * 16 + lane : lane;                  // input channel held by this lane
          const bool live = !DWR_M64 || lane < 16;
          float* dst = stg + ((size_t)slot * 32 + row) * 36;
          if (n == 32) {
            float v[32];
            tmem_ld32(tq + (uint32_t)(col0 + t * 32), v);
            if (live) {
#pragma unroll
              for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          } else {
            float v[16];
            tmem_ld16(tq + (uint32_t)(col0 + t * 16), v);
            if (live) {
#pragma unroll
              for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
#ifndef FK_DW_TRACE_WORKER
      DWTRACE(count, 5);
#endif
      if (lane == 0) { mbar_arrive(tfree); mbar_arrive(sfull); }
    }
  } else if (widx >= 0) {
    // ---- workers: staged accumulators -> weight-norm transform -> bf16 row of X
    long long count = 0;
    int cur_unit = -1;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++count) {
      const int ui = (int)(item / a.n);
      const long long row = a.row_base + item % a.n;
      mbar_wait(sfull, (uint32_t)(count & 1));   // (also orders this thread after the producer's write of the unit descriptor)
      const DwUnit& u = sunit[count & 3];
      if (ui != cur_unit) {   // new unit: refresh the v_hat / coefficient cache (all workers together)
        named_sync(2, 32 * DWR_WORKERS);
        int off = 0;
        for (int k = 0; k < u.nconv; ++k) {
          const DwConv& c = u.conv[k];
          const int rs = c.n + 4, rows = c.ntaps * c.cin;
          if (c.p_g >= 0) {
            const int n4 = c.n / 4;
            for (int e = widx * 32 + lane; e < rows * n4; e += 32 * DWR_WORKERS) {
              const int r = e / n4, q = e - r * n4;
              *reinterpret_cast<float4*>(vh_s + off + r * rs + 4 * q) =
                  __ldg(reinterpret_cast<const float4*>(a.wn_dir + c.w_off + (long long)r * c.n) + q);
            }
            for (int e = widx * 32 + lane; e < 64; e += 32 * DWR_WORKERS)
              cf_s[k * 64 + e] = (e & 31) < c.n ? a.wn_coef[(c.op * 2 + (e >> 5)) * 64 + (e & 31)] : 0.f;
          }
          off += rows * rs;
        }
        named_sync(2, 32 * DWR_WORKERS);
        cur_unit = ui;
      }
      if (widx == 0) DWTRACE(count, 6);
      // pass 1 (lane = output channel): partial dots dW . v_hat over this worker's taps
      float* xd = xdot + (size_t)(count & 1) * 3 * DWR_WORKERS * 32;
      {
        int slot = 0, off = 0;
        for (int k = 0; k < u.nconv; ++k) {
          const DwConv& c = u.conv[k];
          const int rs = c.n + 4;
          if (c.p_g >= 0) {
            float acc = 0.f;
            if (lane < c.n)
              for (int t = 0; t < c.ntaps; ++t) {
                if ((slot + t) % DWR_WORKERS != widx) continue;
                const float* sp = stg + (size_t)(slot + t) * 32 * 36 + lane;
                const float* vp = vh_s + off + (size_t)t * c.cin * rs + lane;
                for (int ci = 0; ci < c.cin; ++ci) acc = fmaf(sp[ci * 36], vp[ci * rs], acc);
              }
            xd[(k * DWR_WORKERS + widx) * 32 + lane] = acc;
          }
          slot += c.ntaps;
          off += c.ntaps * c.cin * rs;
        }
      }
#ifdef FK_DW_TRACE_WORKER
      if (widx == 0) DWTRACE(count, 4);
#endif
      named_sync(3, 32 * DWR_WORKERS);
#ifdef FK_DW_TRACE_WORKER
      if (widx == 0) DWTRACE(count, 5);
#endif
      float* dt = wdot + (size_t)(count & 1) * 3 * 32;
      if (widx < u.nconv) {   // worker k sums the partial dots of convolution k and writes the weight-norm gain gradient
        const int k = widx;
        const DwConv& c = u.conv[k];
        if (c.p_g >= 0) {
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < DWR_WORKERS; ++w) tot += xd[(k * DWR_WORKERS + w) * 32 + lane];
          tot *= a.out_scale;
          dt[k * 32 + lane] = tot;
          if (lane < c.n) a.xrows[xrow_index(a.rld, row, c.p_g + lane)] = __float2bfloat16_rn(tot * cf_s[k * 64 + 32 + lane]);
        }
      }
      named_sync(3, 32 * DWR_WORKERS);
      // pass 2 (lane = input channel): transform and store this worker's taps
      {
        int slot = 0, off = 0;
        for (int k = 0; k < u.nconv; ++k) {
          const DwConv& c = u.conv[k];
          const int rs = c.n + 4;
          const bool wn = c.p_g >= 0;
          const float* dk = dt + k * 32;
          const float* ck = cf_s + k * 64;
          for (int t = 0; t < c.ntaps; ++t) {
            if ((slot + t) % DWR_WORKERS != widx || lane >= c.cin) continue;
            const float* sp = stg + ((size_t)(slot + t) * 32 + lane) * 36;
            const float* vp = vh_s + off + ((size_t)t * c.cin + lane) * rs;
            const long long p0 = c.p_kernel + ((long long)t * c.cin + lane) * c.n;
            for (int h16 = 0; h16 < c.n / 16; ++h16) {
              float w[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(sp + 16 * h16 + 4 * q);
                if (wn) {
                  const float4 h4 = *reinterpret_cast<const float4*>(vp + 16 * h16 + 4 * q);
                  const float4 d4 = *reinterpret_cast<const float4*>(dk + 16 * h16 + 4 * q);
                  const float4 c4 = *reinterpret_cast<const float4*>(ck + 16 * h16 + 4 * q);
                  w[4 * q + 0] = c4.x * (v4.x * a.out_scale - h4.x * d4.x);
                  w[4 * q + 1] = c4.y * (v4.y * a.out_scale - h4.y * d4.y);
                  w[4 * q + 2] = c4.z * (v4.z * a.out_scale - h4.z * d4.z);
                  w[4 * q + 3] = c4.w * (v4.w * a.out_scale - h4.w * d4.w);
                } else {
                  w[4 * q + 0] = v4.x * a.out_scale; w[4 * q + 1] = v4.y * a.out_scale;
                  w[4 * q + 2] = v4.z * a.out_scale; w[4 * q + 3] = v4.w * a.out_scale;
                }
              }
              uint4 q0, q1;
              __nv_bfloat162 b;
              b = __floats2bfloat162_rn(w[0], w[1]); q0.x = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(w[2], w[3]); q0.y = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(w[4], w[5]); q0.z = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(w[6], w[7]); q0.w = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(w[8], w[9]); q1.x = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(w[10], w[11]); q1.y = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(w[12], w[13]); q1.z = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(w[14], w[15]); q1.w = *reinterpret_cast<uint32_t*>(&b);
              uint4* dst = reinterpret_cast<uint4*>(a.xrows + xrow_index(a.rld, row, p0 + 16 * h16));   // a 16-element piece
              dst[0] = q0;                                                                                // never straddles a panel
              dst[1] = q1;
            }
          }
          slot += c.ntaps;
          off += c.ntaps * c.cin * rs;
        }
      }
      __syncwarp();
      if (widx == 0) DWTRACE(count, 7);
      if (lane == 0) mbar_arrive(sfree);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// bias gradients: db[op][co] = sum_{cfg, rows} dz;  grid (nb * 4 dz tiles, splits)
__global__ void tc_db_kernel(const uint8_t* __restrict__ dz, long long n, int nb, const long long* __restrict__ b_off,
                             float* __restrict__ geff) {
  // b_off: [nb][5] bias offsets of V, X, XX, Y, H in the effective-weight gradient
  const int b = blockIdx.x >> 2, tile = blockIdx.x & 3;
  const int ch = threadIdx.x & 31, part = threadIdx.x >> 5;    // 256 threads: 8 row groups x 32 channels
  float acc = 0.f;
  for (long long cfg = blockIdx.y; cfg < n; cfg += gridDim.y) {
    const __half* t = reinterpret_cast<const __half*>(dz + (((size_t)cfg * nb + b) * 4 + tile) * DZ_TILE);
    for (int r = part; r < 128; r += 8) acc += __half2float(t[((ch >> 3) * 128 + r) * 8 + (ch & 7)]);
  }
  __shared__ float red[8][32];
  red[part][ch] = acc;
  __syncthreads();
  if (part == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][ch];
    // dz tiles: 0 dzH -> H bias; 1 dzc -> XX (ch 0..15), Y (16..31); 2 dzX -> X; 3 dzV -> V
    long long off;
    if (tile == 0) off = b_off[b * 5 + 4] + ch;
    else if (tile == 1) off = ch < 16 ? b_off[b * 5 + 2] + ch : b_off[b * 5 + 3] + ch - 16;
    else if (tile == 2) off = b_off[b * 5 + 1] + ch;
    else off = b_off[b * 5 + 0] + ch;
    atomicAdd(geff + off, s);
  }
}

// head: dW_head[ci][k] = sum h_out_last[p][ci] g[p][k], db_head[k] = sum g[p][k]
__global__ void tc_head_dw_kernel(const uint8_t* __restrict__ dump, const float* __restrict__ glog, long long n, int nb, int npos,
                                  int p_first, long long w_off, long long b_off, float* __restrict__ geff) {
  const int ci = threadIdx.x & 31, part = threadIdx.x >> 5;   // 256 threads
  const size_t xt = (size_t)64 * npos, cfg_bytes = (size_t)(nb * NDUMP + 1) * xt;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, bacc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long cfg = blockIdx.x; cfg < n; cfg += gridDim.x) {
    const __half* t = reinterpret_cast<const __half*>(dump + (size_t)cfg * cfg_bytes + (size_t)((nb - 1) * NDUMP + 4) * xt);
    for (int r = part; r < 128; r += 8) {
      const float4 g = *reinterpret_cast<const float4*>(glog + ((size_t)cfg * 128 + r) * 4);
      const float h = __half2float(t[((size_t)(ci >> 3) * npos + p_first + r) * 8 + (ci & 7)]);
      acc[0] += h * g.x; acc[1] += h * g.y; acc[2] += h * g.z; acc[3] += h * g.w;
      if (ci == 0) { bacc[0] += g.x; bacc[1] += g.y; bacc[2] += g.z; bacc[3] += g.w; }
    }
  }
  for (int k = 0; k < 4; ++k) {
    atomicAdd(geff + w_off + ci * 4 + k, acc[k]);
    if (ci == 0) atomicAdd(geff + b_off + k, bacc[k]);
  }
}

// per-sample variants (stochastic reconfiguration): one row of geff per configuration, plain stores
__global__ void tc_db_ps_kernel(const uint8_t* __restrict__ dz, long long n, int nb, const long long* __restrict__ b_off,
                                float* __restrict__ geff, long long row_stride, float out_scale) {
  const int b = blockIdx.x >> 2, tile = blockIdx.x & 3;
  const int ch = threadIdx.x & 31, part = threadIdx.x >> 5;    // 256 threads: 8 row groups x 32 channels
  __shared__ float red[8][32];
  for (long long cfg = blockIdx.y; cfg < n; cfg += gridDim.y) {
    const __half* t = reinterpret_cast<const __half*>(dz + (((size_t)cfg * nb + b) * 4 + tile) * DZ_TILE);
    float acc = 0.f;
    for (int r = part; r < 128; r += 8) acc += __half2float(t[((ch >> 3) * 128 + r) * 8 + (ch & 7)]);
    red[part][ch] = acc;
    __syncthreads();
    if (part == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i][ch];
      long long off;
      if (tile == 0) off = b_off[b * 5 + 4] + ch;
      else if (tile == 1) off = ch < 16 ? b_off[b * 5 + 2] + ch : b_off[b * 5 + 3] + ch - 16;
      else if (tile == 2) off = b_off[b * 5 + 1] + ch;
      else off = b_off[b * 5 + 0] + ch;
      geff[cfg * row_stride + off] = s * out_scale;
    }
    __syncthreads();
  }
}
__global__ void tc_head_dw_ps_kernel(const uint8_t* __restrict__ dump, const float* __restrict__ glog, long long n, int nb, int npos,
                                     int p_first, long long w_off, long long b_off, float* __restrict__ geff, long long row_stride,
                                     float out_scale) {
  const int ci = threadIdx.x & 31, part = threadIdx.x >> 5;   // 256 threads
  const size_t xt = (size_t)64 * npos, cfg_bytes = (size_t)(nb * NDUMP + 1) * xt;
  __shared__ float red[8][32][4], bred[8][4];
  for (long long cfg = blockIdx.x; cfg < n; cfg += gridDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, bacc[4] = {0.f, 0.f, 0.f, 0.f};
    const __half* t = reinterpret_cast<const __half*>(dump + (size_t)cfg * cfg_bytes + (size_t)((nb - 1) * NDUMP + 4) * xt);
    for (int r = part; r < 128; r += 8) {
      const float4 g = *reinterpret_cast<const float4*>(glog + ((size_t)cfg * 128 + r) * 4);
      const float h = __half2float(t[((size_t)(ci >> 3) * npos + p_first + r) * 8 + (ci & 7)]);
      acc[0] += h * g.x; acc[1] += h * g.y; acc[2] += h * g.z; acc[3] += h * g.w;
      if (ci == 0) { bacc[0] += g.x; bacc[1] += g.y; bacc[2] += g.z; bacc[3] += g.w; }
    }
    for (int k = 0; k < 4; ++k) { red[part][ci][k] = acc[k]; if (ci == 0) bred[part][k] = bacc[k]; }
    __syncthreads();
    if (part == 0) {
      for (int k = 0; k < 4; ++k) {
        float sacc = 0.f;
        for (int i = 0; i < 8; ++i) sacc += red[i][ci][k];
        geff[cfg * row_stride + w_off + ci * 4 + k] = sacc * out_scale;
      }
      if (ci < 4) {
        float sb = 0.f;
        for (int i = 0; i < 8; ++i) sb += bred[i][ci];
        geff[cfg * row_stride + b_off + ci] = sb * out_scale;
      }
    }
    __syncthreads();
  }
}
// the same two reductions writing bf16 Jacobian rows (panel-major layout, RAW parameter offsets; biases and the head are
// never weight-normalised)
__global__ void tc_db_rows_kernel(const uint8_t* __restrict__ dz, long long n, int nb, const long long* __restrict__ pb_off,
                                  __nv_bfloat16* __restrict__ xrows, long long rld, long long row_base, float out_scale) {
  const int b = blockIdx.x >> 2, tile = blockIdx.x & 3;
  const int ch = threadIdx.x & 31, part = threadIdx.x >> 5;
  __shared__ float red[8][32];
  for (long long cfg = blockIdx.y; cfg < n; cfg += gridDim.y) {
    const __half* t = reinterpret_cast<const __half*>(dz + (((size_t)cfg * nb + b) * 4 + tile) * DZ_TILE);
    float acc = 0.f;
    for (int r = part; r < 128; r += 8) acc += __half2float(t[((ch >> 3) * 128 + r) * 8 + (ch & 7)]);
    red[part][ch] = acc;
    __syncthreads();
    if (part == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i][ch];
      long long off;
      if (tile == 0) off = pb_off[b * 5 + 4] + ch;
      else if (tile == 1) off = ch < 16 ? pb_off[b * 5 + 2] + ch : pb_off[b * 5 + 3] + ch - 16;
      else if (tile == 2) off = pb_off[b * 5 + 1] + ch;
      else off = pb_off[b * 5 + 0] + ch;
      xrows[xrow_index(rld, row_base + cfg, off)] = __float2bfloat16_rn(s * out_scale);
    }
    __syncthreads();
  }
}
__global__ void tc_head_dw_rows_kernel(const uint8_t* __restrict__ dump, const float* __restrict__ glog, long long n, int nb, int npos,
                                       int p_first, long long pw_off, long long pb_off, __nv_bfloat16* __restrict__ xrows,
                                       long long rld, long long row_base, float out_scale, long long num_params) {
  const int ci = threadIdx.x & 31, part = threadIdx.x >> 5;   // 256 threads
  const size_t xt = (size_t)64 * npos, cfg_bytes = (size_t)(nb * NDUMP + 1) * xt;
  __shared__ float red[8][32][4], bred[8][4];
  for (long long cfg = blockIdx.x; cfg < n; cfg += gridDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, bacc[4] = {0.f, 0.f, 0.f, 0.f};
    const __half* t = reinterpret_cast<const __half*>(dump + (size_t)cfg * cfg_bytes + (size_t)((nb - 1) * NDUMP + 4) * xt);
    for (int r = part; r < 128; r += 8) {
      const float4 g = *reinterpret_cast<const float4*>(glog + ((size_t)cfg * 128 + r) * 4);
      const float h = __half2float(t[((size_t)(ci >> 3) * npos + p_first + r) * 8 + (ci & 7)]);
      acc[0] += h * g.x; acc[1] += h * g.y; acc[2] += h * g.z; acc[3] += h * g.w;
      if (ci == 0) { bacc[0] += g.x; bacc[1] += g.y; bacc[2] += g.z; bacc[3] += g.w; }
    }
    for (int k = 0; k < 4; ++k) { red[part][ci][k] = acc[k]; if (ci == 0) bred[part][k] = bacc[k]; }
    __syncthreads();
    if (part == 0) {
      for (int k = 0; k < 4; ++k) {
        float sacc = 0.f;
        for (int i = 0; i < 8; ++i) sacc += red[i][ci][k];
        xrows[xrow_index(rld, row_base + cfg, pw_off + ci * 4 + k)] = __float2bfloat16_rn(sacc * out_scale);
      }
      if (ci < 4) {
        float sb = 0.f;
        for (int i = 0; i < 8; ++i) sb += bred[i][ci];
        xrows[xrow_index(rld, row_base + cfg, pb_off + ci)] = __float2bfloat16_rn(sb * out_scale);
      }
      // zero padding of the last panel (columns num_params .. next multiple of 64)
      const long long pad0 = num_params + ci;
      if (pad0 < ((num_params + 63) & ~63LL)) xrows[xrow_index(rld, row_base + cfg, pad0)] = __float2bfloat16_rn(0.f);
      if (pad0 + 32 < ((num_params + 63) & ~63LL)) xrows[xrow_index(rld, row_base + cfg, pad0 + 32)] = __float2bfloat16_rn(0.f);
    }
    __syncthreads();
  }
}
__global__ void tc_const_coef_kernel(long long n, float re, float im, float* __restrict__ coef) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  coef[2 * i] = re;
  coef[2 * i + 1] = im;
}

// power-of-two loss scale chosen on the device: max |2 y| * scale <= 128 (head gradients in the fp16 normal range)
__global__ void tc_absmax_kernel(const float* __restrict__ y, long long n2, unsigned int* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) atomicMax(out, __float_as_uint(fabsf(y[i])));
}
__global__ void tc_setscale_kernel(const unsigned int* __restrict__ absmax, float* __restrict__ scale) {
  const float m = __uint_as_float(*absmax);
  float s = 1.f;
  if (m > 0.f && isfinite(m)) s = exp2f(fminf(fmaxf(floorf(log2f(64.f / m)), -60.f), 60.f));
  scale[0] = s;
  scale[1] = 1.f / s;
}
__global__ void tc_coef_kernel(const float* __restrict__ y, long long n, const float* __restrict__ scale, float* __restrict__ coef) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  coef[2 * i] = 2.f * y[2 * i] * scale[0];         // d L / d Re log psi
  coef[2 * i + 1] = -2.f * y[2 * i + 1] * scale[0];  // d L / d Im log psi
}
__global__ void tc_scale_kernel(float* __restrict__ p, long long n, const float* __restrict__ scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= scale[1];
}

// ================================================================================================================
// host
// ================================================================================================================
int tc_grad_supported(const fk_net* net) {
  if (!tc_supported(net)) return 0;
  TcPublicGeometry g;
  if (tc_public_geometry(net, &g)) return 0;
  return g.T == 1 ? 1 : 0;
}

static size_t bwd_units_offset(int nb) { return ((size_t)nb * IMGB_BYTES + 255) / 256 * 256; }
static size_t bwd_boff_offset(int nb) { return bwd_units_offset(nb) + ((size_t)2 * nb * sizeof(DwUnit) + 255) / 256 * 256; }
static size_t bwd_pboff_offset(int nb) { return bwd_boff_offset(nb) + ((size_t)nb * 5 * sizeof(long long) + 255) / 256 * 256; }
static size_t bwd_pack_offset(int nb) { return bwd_pboff_offset(nb) + ((size_t)nb * 5 * sizeof(long long) + 255) / 256 * 256; }

int tc_grad_prepare(fk_net* net) {
  if (!tc_grad_supported(net)) return 0;
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  const int nb = g.nb;
  {
    FK_CHECK_CUDA(cudaMalloc(&net->d_tc_bwd, bwd_pack_offset(nb) + sizeof(BwdPackDesc) * nb));
    std::vector<BwdPackDesc> pd(nb);
    std::vector<long long> boff(nb * 5), pboff(nb * 5);
    std::vector<DwUnit> units(2 * nb);
    auto res2 = [&](int b) { return b >= 2 && b % 2 == 0 && b != nb - 1; };
    for (int b = 0; b < nb; ++b) {
      const ConvOp* o = &net->ops[5 * b];
      for (int r = 0; r < 5; ++r) { pd[b].w[r] = o[r].w_off; boff[b * 5 + r] = o[r].b_off; pboff[b * 5 + r] = o[r].p_bias; }
      pd[b].w_head = net->ops.back().w_off;
      pd[b].cin = b == 0 ? 1 : 32;
      const bool last = b == nb - 1;
      const int in_tile = nb * NDUMP;
      const int vin = b == 0 ? in_tile : (res2(b - 1) ? (b - 1) * NDUMP + 2 : (b - 1) * NDUMP + 1);
      const int hin = b == 0 ? in_tile : (b - 1) * NDUMP + 4;
      DwUnit& A = units[2 * b];        // pass A: H, XX, Y
      A.b = b; A.nconv = 3;
      A.conv[0] = DwConv{b * NDUMP + 3, 0, 0, 9, 32, {0}, 0, 32, o[4].w_off, o[4].p_kernel, o[4].p_g, 5 * b + 4};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A.conv[0].off[i * 3 + j] = g.p_first + (i - 2) * g.P + (j - 2);
      A.conv[1] = DwConv{b * NDUMP + 0, 1, 0, 1, 16, {0}, 288, 32, o[2].w_off, o[2].p_kernel, o[2].p_g, 5 * b + 2};
      A.conv[1].off[0] = g.p_first + (last ? -1 : 0);
      A.conv[2] = DwConv{b * NDUMP + 1, 1, 2, 1, 16, {0}, 304, 32, o[3].w_off, o[3].p_kernel, o[3].p_g, 5 * b + 3};
      A.conv[2].off[0] = g.p_first - g.P;
      DwUnit& Bu = units[2 * b + 1];   // pass B: V, X
      Bu.b = b; Bu.nconv = 2;
      Bu.conv[0] = DwConv{vin, 3, 0, 9, 32, {0}, 0, pd[b].cin, o[0].w_off, o[0].p_kernel, o[0].p_g, 5 * b + 0};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Bu.conv[0].off[i * 3 + j] = g.p_first + (i - 2) * g.P + (j - 1);
      Bu.conv[1] = DwConv{hin, 2, 0, 3, 32, {0}, 288, pd[b].cin, o[1].w_off, o[1].p_kernel, o[1].p_g, 5 * b + 1};
      for (int j = 0; j < 3; ++j) Bu.conv[1].off[j] = g.p_first + (j - 2);
    }
    uint8_t* base = (uint8_t*)net->d_tc_bwd;
    FK_CHECK_CUDA(cudaMemcpy(base + bwd_units_offset(nb), units.data(), sizeof(DwUnit) * units.size(), cudaMemcpyHostToDevice));
    FK_CHECK_CUDA(cudaMemcpy(base + bwd_boff_offset(nb), boff.data(), sizeof(long long) * boff.size(), cudaMemcpyHostToDevice));
    FK_CHECK_CUDA(cudaMemcpy(base + bwd_pboff_offset(nb), pboff.data(), sizeof(long long) * pboff.size(), cudaMemcpyHostToDevice));
    FK_CHECK_CUDA(cudaMemcpy(base + bwd_pack_offset(nb), pd.data(), sizeof(BwdPackDesc) * nb, cudaMemcpyHostToDevice));
  }
  return 0;
}

int tc_grad_pack_weights(fk_net* net, cudaStream_t s) {
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  const int nb = g.nb;
  FK_REQUIRE(net->d_tc_bwd, "tc_grad_pack_weights: the machine was created without the tensor-core gradient tables");
  tc_pack_bwd_kernel<<<nb, 256, 0, s>>>(net->d_weff, reinterpret_cast<const BwdPackDesc*>((uint8_t*)net->d_tc_bwd + bwd_pack_offset(nb)),
                                       (uint8_t*)net->d_tc_bwd);
  FK_CHECK_LAUNCH();
  return 0;
}

struct GradLayout { size_t dump, mask, logits, dz, glog, coef, geff, lp, scale, total; int64_t chunk; };

static GradLayout grad_layout(const fk_net* net, int64_t chunk) {
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  GradLayout L;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  L.scale = o; o = al(o + 64);
  L.geff = o; o = al(o + sizeof(float) * net->num_eff);
  L.dump = o; o = al(o + (size_t)chunk * (g.nb * NDUMP + 1) * 64 * g.npos);
  L.mask = o; o = al(o + (size_t)chunk * g.nb * NDUMP * 128 * 4);
  L.logits = o; o = al(o + (size_t)chunk * 128 * 16);
  L.dz = o; o = al(o + (size_t)chunk * g.nb * 4 * DZ_TILE);
  L.glog = o; o = al(o + (size_t)chunk * 128 * 16);
  L.coef = o; o = al(o + (size_t)chunk * 8);
  L.lp = o; o = al(o + (size_t)chunk * 8);
  L.total = o; L.chunk = chunk;
  return L;
}

int64_t tc_grad_workspace_bytes(const fk_net* net, int64_t B) {
  return (int64_t)grad_layout(net, std::max<int64_t>(1, std::min<int64_t>(B, 2048))).total;
}

int tc_grad_weighted(fk_net* net, const int8_t* sigma, const float* y, int64_t B, float* grad_out, void* ws, int64_t ws_bytes,
                     cudaStream_t s) {
  FK_REQUIRE(net->params_set && net->d_tc_bwd, "tensor-core gradient weights were never packed (fk_net_set_params)");
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  int64_t chunk = std::min<int64_t>(B, 2048);
  while (chunk > 1 && (int64_t)grad_layout(net, chunk).total > ws_bytes) chunk /= 2;
  const GradLayout L = grad_layout(net, chunk);
  FK_REQUIRE((int64_t)L.total <= ws_bytes, "tensor-core gradient: workspace too small (%lld < %zu bytes)", (long long)ws_bytes, L.total);
  uint8_t* base = (uint8_t*)ws;
  float* geff = (float*)(base + L.geff);
  FK_CHECK_CUDA(cudaMemsetAsync(geff, 0, sizeof(float) * net->num_eff, s));
  unsigned int* absmax = (unsigned int*)(base + L.scale);
  float* scale = (float*)(base + L.scale) + 4;
  FK_CHECK_CUDA(cudaMemsetAsync(absmax, 0, 16, s));
  tc_absmax_kernel<<<(unsigned)((2 * B + 255) / 256), 256, 0, s>>>(y, 2 * B, absmax);
  FK_CHECK_LAUNCH();
  tc_setscale_kernel<<<1, 1, 0, s>>>(absmax, scale);
  FK_CHECK_LAUNCH();
  const int nb = g.nb;
  const int npos_g = ((g.p_first + 128 + 2 * g.P + 4) + 7) / 8 * 8;
  int dev = 0, sms = 148;
  FK_CHECK_CUDA(cudaGetDevice(&dev));
  FK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t bwd_smem = 2 * (size_t)IMGB_BYTES + (size_t)BWD_NP * 4 * 64 * npos_g + 256;
  const size_t dw_smem = (size_t)DW_STAGES * 3 * (64 * g.npos + DZ_TILE) + (size_t)16 * g.npos * 16 + DW_TAIL_BYTES;
  FK_REQUIRE(bwd_smem <= 227 * 1024 && dw_smem <= 227 * 1024, "tensor-core gradient: lattice too large for shared memory");
  FK_CHECK_CUDA(cudaFuncSetAttribute(tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
  FK_CHECK_CUDA(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem));
  const uint8_t* wb = (const uint8_t*)net->d_tc_bwd;
  for (int64_t i = 0; i < B; i += chunk) {
    const int64_t m = std::min(chunk, B - i);
    const int8_t* sg = sigma + i * net->sites;
    tc_coef_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(y + 2 * i, m, scale, (float*)(base + L.coef));
    FK_CHECK_LAUNCH();
    if (tc_forward_launch(net, sg, m, (float*)(base + L.lp), base + L.dump, (uint32_t*)(base + L.mask), (float*)(base + L.logits), s)) return 1;
    BwdArgs ba;
    ba.images = wb; ba.mask = (const uint32_t*)(base + L.mask); ba.logits = (const float*)(base + L.logits);
    ba.sigma = sg; ba.coef = (const float*)(base + L.coef); ba.dz = base + L.dz; ba.glog = (float*)(base + L.glog);
    ba.n = m; ba.H = net->H; ba.W = net->W; ba.P = g.P; ba.nb = nb; ba.npos_g = npos_g; ba.p_first = g.p_first;
    const long long groups = (m + BWD_NP - 1) / BWD_NP;
    tc_backward_kernel<<<(unsigned)std::min<long long>(groups, sms), BWD_NP * 128 + 32, bwd_smem, s>>>(ba);
    FK_CHECK_LAUNCH();
    DwArgs2 da;
    da.dump = base + L.dump; da.dz = base + L.dz; da.units = reinterpret_cast<const DwUnit*>(wb + bwd_units_offset(nb));
    da.geff = geff; da.n = m; da.nb = nb; da.npos = g.npos; da.num_units = 2 * nb;
    da.cfg_chunk = (int)std::max<int64_t>(1, (m + 7) / 8);
    da.row_stride = 0; da.out_scale = 1.f;
    da.xrows = nullptr; da.rld = 0; da.row_base = 0; da.wn_dir = nullptr; da.wn_coef = nullptr; da.stages = DW_STAGES;
    const int items = da.num_units * (int)((m + da.cfg_chunk - 1) / da.cfg_chunk);
    tc_dw_kernel<<<(unsigned)std::min(items, sms), 128, dw_smem, s>>>(da);
    FK_CHECK_LAUNCH();
    tc_db_kernel<<<dim3((unsigned)(nb * 4), 8), 256, 0, s>>>(base + L.dz, m, nb, reinterpret_cast<const long long*>(wb + bwd_boff_offset(nb)), geff);
    FK_CHECK_LAUNCH();
    tc_head_dw_kernel<<<64, 256, 0, s>>>(base + L.dump, (const float*)(base + L.glog), m, nb, g.npos, g.p_first,
                                         net->ops.back().w_off, net->ops.back().b_off, geff);
    FK_CHECK_LAUNCH();
  }
  tc_scale_kernel<<<(unsigned)((net->num_eff + 255) / 256), 256, 0, s>>>(geff, net->num_eff, scale);
  FK_CHECK_LAUNCH();
  if (grad_transform_launch(net, geff, grad_out, s)) return 1;
  return 0;
}

// ---- per-sample Jacobians on the tensor cores: O_re[b] = d Re log psi(sigma_b) / d theta, O_im[b] = d Im log psi / d theta ----
struct PsLayout { size_t geff, dump, mask, logits, dz, glog, coef, lp, total; };
static PsLayout ps_layout(const fk_net* net, int64_t chunk) {
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  PsLayout L;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  L.geff = o; o = al(o + sizeof(float) * (size_t)net->num_eff * chunk);
  L.dump = o; o = al(o + (size_t)chunk * (g.nb * NDUMP + 1) * 64 * g.npos);
  L.mask = o; o = al(o + (size_t)chunk * g.nb * NDUMP * 128 * 4);
  L.logits = o; o = al(o + (size_t)chunk * 128 * 16);
  L.dz = o; o = al(o + (size_t)chunk * g.nb * 4 * DZ_TILE);
  L.glog = o; o = al(o + (size_t)chunk * 128 * 16);
  L.coef = o; o = al(o + (size_t)chunk * 8);
  L.lp = o; o = al(o + (size_t)chunk * 8);
  L.total = o;
  return L;
}

int64_t tc_grad_per_sample_workspace_bytes(const fk_net* net, int64_t B) {
  return (int64_t)ps_layout(net, std::max<int64_t>(1, std::min<int64_t>(B, 512))).total;
}

int tc_grad_per_sample(fk_net* net, const int8_t* sigma, int64_t B, float* O_re, float* O_im, void* ws, int64_t ws_bytes,
                       cudaStream_t s) {
  FK_REQUIRE(net->params_set && net->d_tc_bwd, "tensor-core gradient weights were never packed (fk_net_set_params)");
  FK_REQUIRE(net->num_eff % 4 == 0, "tensor-core per-sample gradient: effective-weight rows must be 16-byte aligned");
  for (const ConvOp& op : net->ops)
    FK_REQUIRE(op.w_off % 4 == 0, "tensor-core per-sample gradient: weight offsets must be 16-byte aligned");
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  int64_t chunk = std::min<int64_t>(B, 512);
  while (chunk > 1 && (int64_t)ps_layout(net, chunk).total > ws_bytes) chunk /= 2;
  const PsLayout L = ps_layout(net, chunk);
  FK_REQUIRE((int64_t)L.total <= ws_bytes, "tensor-core per-sample gradient: workspace too small (%lld < %zu bytes)",
             (long long)ws_bytes, L.total);
  uint8_t* base = (uint8_t*)ws;
  float* geff = (float*)(base + L.geff);
  const int nb = g.nb;
  const int npos_g = ((g.p_first + 128 + 2 * g.P + 4) + 7) / 8 * 8;
  int dev = 0, sms = 148;
  FK_CHECK_CUDA(cudaGetDevice(&dev));
  FK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t bwd_smem = 2 * (size_t)IMGB_BYTES + (size_t)BWD_NP * 4 * 64 * npos_g + 256;
  const size_t dw_smem = (size_t)DW_STAGES * 3 * (64 * g.npos + DZ_TILE) + (size_t)16 * g.npos * 16 + DW_TAIL_BYTES;
  FK_REQUIRE(bwd_smem <= 227 * 1024 && dw_smem <= 227 * 1024, "tensor-core gradient: lattice too large for shared memory");
  FK_CHECK_CUDA(cudaFuncSetAttribute(tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
  FK_CHECK_CUDA(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem));
  const uint8_t* wb = (const uint8_t*)net->d_tc_bwd;
  const float seed_scale = 64.f;   // power of two: unit seeds x 64 keep the head gradients in the fp16 normal range
  for (int64_t i = 0; i < B; i += chunk) {
    const int64_t m = std::min(chunk, B - i);
    const int8_t* sg = sigma + i * net->sites;
    if (tc_forward_launch(net, sg, m, (float*)(base + L.lp), base + L.dump, (uint32_t*)(base + L.mask), (float*)(base + L.logits), s)) return 1;
    for (int pass = 0; pass < (O_im ? 2 : 1); ++pass) {
      // coef = (dL/dRe log psi, dL/dIm log psi): (1, 0) gives d Re log psi / d theta, (0, 1) gives d Im log psi / d theta
      tc_const_coef_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(m, pass == 0 ? seed_scale : 0.f, pass == 0 ? 0.f : seed_scale,
                                                                        (float*)(base + L.coef));
      FK_CHECK_LAUNCH();
      BwdArgs ba;
      ba.images = wb; ba.mask = (const uint32_t*)(base + L.mask); ba.logits = (const float*)(base + L.logits);
      ba.sigma = sg; ba.coef = (const float*)(base + L.coef); ba.dz = base + L.dz; ba.glog = (float*)(base + L.glog);
      ba.n = m; ba.H = net->H; ba.W = net->W; ba.P = g.P; ba.nb = nb; ba.npos_g = npos_g; ba.p_first = g.p_first;
      const long long groups = (m + BWD_NP - 1) / BWD_NP;
      tc_backward_kernel<<<(unsigned)std::min<long long>(groups, sms), BWD_NP * 128 + 32, bwd_smem, s>>>(ba);
      FK_CHECK_LAUNCH();
      DwArgs2 da;
      da.dump = base + L.dump; da.dz = base + L.dz; da.units = reinterpret_cast<const DwUnit*>(wb + bwd_units_offset(nb));
      da.geff = geff; da.n = m; da.nb = nb; da.npos = g.npos; da.num_units = 2 * nb;
      da.cfg_chunk = 1; da.row_stride = net->num_eff; da.out_scale = 1.f / seed_scale;
      da.xrows = nullptr; da.rld = 0; da.row_base = 0; da.wn_dir = nullptr; da.wn_coef = nullptr; da.stages = DW_STAGES;
      const long long items = (long long)da.num_units * m;
      tc_dw_kernel<<<(unsigned)std::min<long long>(items, sms), 128, dw_smem, s>>>(da);
      FK_CHECK_LAUNCH();
      tc_db_ps_kernel<<<dim3((unsigned)(nb * 4), (unsigned)std::min<int64_t>(m, 64)), 256, 0, s>>>(
          base + L.dz, m, nb, reinterpret_cast<const long long*>(wb + bwd_boff_offset(nb)), geff, net->num_eff, 1.f / seed_scale);
      FK_CHECK_LAUNCH();
      tc_head_dw_ps_kernel<<<(unsigned)std::min<int64_t>(m, 1024), 256, 0, s>>>(
          base + L.dump, (const float*)(base + L.glog), m, nb, g.npos, g.p_first, net->ops.back().w_off, net->ops.back().b_off, geff,
          net->num_eff, 1.f / seed_scale);
      FK_CHECK_LAUNCH();
      if (grad_transform_rows_launch(net, geff, (pass == 0 ? O_re : O_im) + i * net->num_params, m, s)) return 1;
    }
  }
  return 0;
}

// ---- Jacobian rows straight into the Gram operand -------------------------------------------------------------------
// X: bf16, panel-major [ceil(P / 64)][rld rows][64]; sample b writes row row_re + b (d Re log psi / d theta) and, when
// row_im >= 0, row row_im + b (d Im log psi / d theta), in RAW parameters (weight-norm transform fused into the flush).
struct RowsLayout { size_t dump, mask, logits, dz, glog, coef, lp, total; };
static RowsLayout rows_layout(const fk_net* net, int64_t chunk) {
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  RowsLayout L;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  L.dump = o; o = al(o + (size_t)chunk * (g.nb * NDUMP + 1) * 64 * g.npos);
  L.mask = o; o = al(o + (size_t)chunk * g.nb * NDUMP * 128 * 4);
  L.logits = o; o = al(o + (size_t)chunk * 128 * 16);
  L.dz = o; o = al(o + (size_t)chunk * g.nb * 4 * DZ_TILE);
  L.glog = o; o = al(o + (size_t)chunk * 128 * 16);
  L.coef = o; o = al(o + (size_t)chunk * 8);
  L.lp = o; o = al(o + (size_t)chunk * 8);
  L.total = o;
  return L;
}

int64_t tc_jacobian_rows_workspace_bytes(const fk_net* net, int64_t B) {
  return (int64_t)rows_layout(net, std::max<int64_t>(1, std::min<int64_t>(B, 1024))).total;
}

int tc_jacobian_rows(fk_net* net, const int8_t* sigma, int64_t B, void* X, int64_t rld, int64_t row_re, int64_t row_im,
                     void* ws, int64_t ws_bytes, cudaStream_t s) {
  FK_REQUIRE(net->params_set && net->d_tc_bwd && net->d_wn_dir, "tensor-core gradient weights were never packed (fk_net_set_params)");
  for (const ConvOp& op : net->ops)
    FK_REQUIRE(op.p_kernel % 16 == 0 && op.w_off % 4 == 0, "tc_jacobian_rows: kernel offsets must be multiples of 16 parameters");
  TcPublicGeometry g;
  tc_public_geometry(net, &g);
  int64_t chunk = std::min<int64_t>(B, 1024);
  while (chunk > 1 && (int64_t)rows_layout(net, chunk).total > ws_bytes) chunk /= 2;
  const RowsLayout L = rows_layout(net, chunk);
  FK_REQUIRE((int64_t)L.total <= ws_bytes, "tc_jacobian_rows: workspace too small (%lld < %zu bytes)", (long long)ws_bytes, L.total);
  uint8_t* base = (uint8_t*)ws;
  const int nb = g.nb;
  const int npos_g = ((g.p_first + 128 + 2 * g.P + 4) + 7) / 8 * 8;
  int dev = 0, sms = 148;
  FK_CHECK_CUDA(cudaGetDevice(&dev));
  FK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t bwd_smem = 2 * (size_t)IMGB_BYTES + (size_t)BWD_NP * 4 * 64 * npos_g + 256;
  const size_t dw_smem = (size_t)DW_ROWS_STAGES * 3 * (64 * g.npos + DZ_TILE) + DWR_TAIL_BYTES;
  FK_REQUIRE(bwd_smem <= 227 * 1024 && dw_smem <= 227 * 1024, "tensor-core gradient: lattice too large for shared memory");
  // the garbage M rows of the last A tile read up to 16 channel groups past its start: that must stay inside the allocation
  FK_REQUIRE((size_t)DW_ROWS_STAGES * 3 * (64 * g.npos + DZ_TILE) - 3 * DZ_TILE - 64 * g.npos + (size_t)16 * g.npos * 16 + 4096 <= dw_smem,
             "tc_jacobian_rows: operand ring + staging layout does not cover the slack rows");
  FK_CHECK_CUDA(cudaFuncSetAttribute(tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem));
  FK_CHECK_CUDA(cudaFuncSetAttribute(tc_dw_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem));
  const uint8_t* wb = (const uint8_t*)net->d_tc_bwd;
  const float seed_scale = 64.f;
  __nv_bfloat16* xr = (__nv_bfloat16*)X;
  for (int64_t i = 0; i < B; i += chunk) {
    const int64_t m = std::min(chunk, B - i);
    const int8_t* sg = sigma + i * net->sites;
    if (tc_forward_launch(net, sg, m, (float*)(base + L.lp), base + L.dump, (uint32_t*)(base + L.mask), (float*)(base + L.logits), s)) return 1;
    for (int pass = 0; pass < (row_im >= 0 ? 2 : 1); ++pass) {
      const long long row_base = (pass == 0 ? row_re : row_im) + i;
      tc_const_coef_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(m, pass == 0 ? seed_scale : 0.f, pass == 0 ? 0.f : seed_scale,
                                                                        (float*)(base + L.coef));
      FK_CHECK_LAUNCH();
      BwdArgs ba;
      ba.images = wb; ba.mask = (const uint32_t*)(base + L.mask); ba.logits = (const float*)(base + L.logits);
      ba.sigma = sg; ba.coef = (const float*)(base + L.coef); ba.dz = base + L.dz; ba.glog = (float*)(base + L.glog);
      ba.n = m; ba.H = net->H; ba.W = net->W; ba.P = g.P; ba.nb = nb; ba.npos_g = npos_g; ba.p_first = g.p_first;
      const long long groups = (m + BWD_NP - 1) / BWD_NP;
      tc_backward_kernel<<<(unsigned)std::min<long long>(groups, sms), BWD_NP * 128 + 32, bwd_smem, s>>>(ba);
      FK_CHECK_LAUNCH();
      DwArgs2 da;
      da.dump = base + L.dump; da.dz = base + L.dz; da.units = reinterpret_cast<const DwUnit*>(wb + bwd_units_offset(nb));
      da.geff = nullptr; da.n = m; da.nb = nb; da.npos = g.npos; da.num_units = 2 * nb;
      da.cfg_chunk = 1; da.row_stride = 1; da.out_scale = 1.f / seed_scale;
      da.xrows = xr; da.rld = rld; da.row_base = row_base; da.wn_dir = net->d_wn_dir; da.wn_coef = net->d_wn_coef;
      da.stages = DW_ROWS_STAGES;
      const long long items = (long long)da.num_units * m;
      tc_dw_rows_kernel<<<(unsigned)std::min<long long>(items, sms), DWR_THREADS, dw_smem, s>>>(da);
      FK_CHECK_LAUNCH();
      tc_db_rows_kernel<<<dim3((unsigned)(nb * 4), (unsigned)std::min<int64_t>(m, 64)), 256, 0, s>>>(
          base + L.dz, m, nb, reinterpret_cast<const long long*>(wb + bwd_pboff_offset(nb)), xr, rld, row_base, 1.f / seed_scale);
      FK_CHECK_LAUNCH();
      tc_head_dw_rows_kernel<<<(unsigned)std::min<int64_t>(m, 1024), 256, 0, s>>>(
          base + L.dump, (const float*)(base + L.glog), m, nb, g.npos, g.p_first, net->ops.back().p_kernel, net->ops.back().p_bias, xr,
          rld, row_base, 1.f / seed_scale, net->num_params);
      FK_CHECK_LAUNCH();
    }
  }
  return 0;
}

}  // namespace fk

#ifdef FK_DW_TRACE
extern "C" int fk_dw_trace_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, fk::fk_dw_trace_buf, sizeof(long long) * 64 * 8);
}
#endif
