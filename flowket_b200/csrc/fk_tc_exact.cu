// Tensor-core engine at the contract accuracy ("tc-exact"): the same fused ConvNetAutoregressive2D forward as fk_tc.cu
// (machines/conv_net_autoregressive_2D.py:24-74, machines/abstract_machine.py:31-57), with every operand split into two
// fp16 numbers, x = hi + lo (22 significant bits), and every product formed from three tensor-core passes
//     A B  ~=  A_hi B_hi + A_hi B_lo + A_lo B_hi          (the lo x lo term is below 2^-22)
// packed into TWO tcgen05.mma per (tap, k-step):
//     MMA 1:  A_hi  x  [B_hi | B_lo]   N = 2n   -> accumulator columns [c, c+n) = hi*hi, [c+n, c+2n) = hi*lo
//     MMA 2:  A_lo  x   B_hi           N = n    -> accumulated into the small-term columns [c+n, c+2n)
// so that the A tile (the dominant shared-memory operand stream) is fetched twice, not three times, and the large and
// the small partial sums never share an fp32 accumulator (the tensor core truncates once per MMA: keeping hi*hi alone
// bounds that bias by the 18 MMAs of one convolution).  The epilogue adds the two column groups in fp32, applies
// bias (via a ones x [b_hi, b_lo] MMA) / residual / relu exactly as the fp16 engine does, and writes the next layer's
// operand as an (hi, lo) pair of fp16 tiles.  Weights and biases are pre-scaled by 2^8 so that their lo parts stay in
// the fp16 normal range; the epilogue multiplies by 2^-8 (exact).
//
// Shared memory: 2 pipelines x 3 activation slots x (hi + lo tile) + ONE weight image per block, streamed in two
// chunks (A: the phase-1 convolutions, B: phases 2-4) so that chunk A of block b+1 lands while phases 2-3 of block b
// run and chunk B while phase 1 of block b+1 runs -- double buffering at a single image's footprint.  The horizontal
// stack's residual input lives in fp32 registers.  Lattices that fit one 128-row tile (up to 10 x 10).
#include <algorithm>

#include "fk_net.cuh"
#include "fk_tc_common.cuh"

namespace fk {

struct TcBlockDesc {   // same wiring record as fk_tc.cu
  int8_t in_v, in_h, out_a, out_r, res_v, x1, c, out_h, res_h, last, save_h, pad1;
};

// ---- weight image (bytes).  A B tile for one (tap, k-step) is [k group 0..1][2n rows: n hi + n lo][8 el] fp16.
constexpr int XA_X = 0;            // 3 taps x 2 k-steps x 2048 B (n = 32)
constexpr int XA_V = 12288;        // 9 x 2 x 2048
constexpr int XA_BT_X = 49152;     // bias tile: 32 rows x [hi, lo, 0 x 6]
constexpr int XA_BT_V = 49664;
constexpr int XA_BYTES = 50176;
constexpr int XB_XX = 0;           // 2 k-steps x 1024 B (n = 16)
constexpr int XB_Y = 2048;
constexpr int XB_H = 4096;         // 9 x 2 x 2048
constexpr int XB_HEAD = 40960;     // 2 x 1024 (n = 16, 4 real columns)
constexpr int XB_BT_XX = 43008;    // 16 rows
constexpr int XB_BT_Y = 43264;
constexpr int XB_BT_H = 43520;     // 32 rows
constexpr int XB_BT_HEAD = 44032;  // 16 rows
constexpr int XB_BYTES = 44288;
constexpr int XIMG_BYTES = XA_BYTES + XB_BYTES;
constexpr float X_WSCALE = 256.f, X_WINV = 1.f / 256.f;
constexpr int X_NP = 2, X_SLOTS = 3, X_ISSUERS = 3;
constexpr int X_EPI = 256;   // epilogue threads per pipeline: TWO per lattice position, 16 channels each (see the kernel header)

struct XPackDesc {
  long long w[5], b[5];  // V, X, XX, Y, H
  long long w_head, b_head;
  int cin;
};

__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

__global__ void tcx_pack_kernel(const float* __restrict__ weff, const XPackDesc* __restrict__ pd, uint8_t* __restrict__ images) {
  const XPackDesc d = pd[blockIdx.x];
  uint8_t* img = images + (size_t)blockIdx.x * XIMG_BYTES;
  const int reg_off[6] = {XA_V, XA_X, XA_BYTES + XB_XX, XA_BYTES + XB_Y, XA_BYTES + XB_H, XA_BYTES + XB_HEAD};
  const int reg_taps[6] = {9, 3, 1, 1, 9, 1};
  const int reg_n[6] = {32, 32, 16, 16, 32, 16};
  for (int r = 0; r < 6; ++r) {
    const int N = reg_n[r], taps = reg_taps[r];
    const int cin = (r == 0 || r == 1) ? d.cin : 32;
    const int nreal = (r == 5) ? 4 : N;
    const long long woff = (r == 5) ? d.w_head : d.w[r];
    __half* w16 = reinterpret_cast<__half*>(img + reg_off[r]);
    const int total = taps * 2 * 2 * N * 8;   // (tap, kstep, kgroup, n, e): one (hi, lo) pair each
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int el = e & 7;
      const int n = (e >> 3) % N;
      const int g = ((e >> 3) / N) & 1;
      const int ks = ((e >> 3) / N / 2) & 1;
      const int tap = (e >> 3) / N / 4;
      const int ci = ks * 16 + g * 8 + el;
      float v = 0.f;
      if (ci < cin && n < nreal) v = weff[woff + ((long long)tap * cin + ci) * nreal + n] * X_WSCALE;
      __half hi, lo;
      split_h(v, hi, lo);
      const int tile = (tap * 2 + ks) * (2 * 2 * N * 8);
      w16[tile + (g * 2 * N + n) * 8 + el] = hi;
      w16[tile + (g * 2 * N + N + n) * 8 + el] = lo;
    }
  }
  // bias tiles: per output channel one 16-byte row [hi, lo, 0 x 6] of 256 * bias
  const int bt_off[6] = {XA_BT_V, XA_BT_X, XA_BYTES + XB_BT_XX, XA_BYTES + XB_BT_Y, XA_BYTES + XB_BT_H, XA_BYTES + XB_BT_HEAD};
  for (int r = 0; r < 6; ++r) {
    const int N = reg_n[r], nreal = (r == 5) ? 4 : N;
    const long long boff = (r == 5) ? d.b_head : d.b[r];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const float bv = i < nreal ? weff[boff + i] * X_WSCALE : 0.f;
      __half hi, lo;
      split_h(bv, hi, lo);
      __half* row = reinterpret_cast<__half*>(img + bt_off[r] + 16 * i);
      row[0] = hi; row[1] = lo;
      for (int k = 2; k < 8; ++k) row[k] = __float2half_rn(0.f);
    }
  }
}

struct TcxArgs {
  const uint8_t* images;
  const TcBlockDesc* desc;
  const int8_t* sigma;
  float* out;
  long long n;
  int H, W, P, nb, npos, p_first, cst_off;
  TcWork wk;   // local-energy work list (see fk_net.cuh)
  TcxPrefix px;   // prefix reuse (see below); all-null = off
};

// two 32-column TMEM loads, one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr, float (&a)[32], float (&b)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]), "=f"(a[8]),
        "=f"(a[9]), "=f"(a[10]), "=f"(a[11]), "=f"(a[12]), "=f"(a[13]), "=f"(a[14]), "=f"(a[15]), "=f"(a[16]),
        "=f"(a[17]), "=f"(a[18]), "=f"(a[19]), "=f"(a[20]), "=f"(a[21]), "=f"(a[22]), "=f"(a[23]), "=f"(a[24]),
        "=f"(a[25]), "=f"(a[26]), "=f"(a[27]), "=f"(a[28]), "=f"(a[29]), "=f"(a[30]), "=f"(a[31])
      : "r"(taddr)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3]), "=f"(b[4]), "=f"(b[5]), "=f"(b[6]), "=f"(b[7]), "=f"(b[8]),
        "=f"(b[9]), "=f"(b[10]), "=f"(b[11]), "=f"(b[12]), "=f"(b[13]), "=f"(b[14]), "=f"(b[15]), "=f"(b[16]),
        "=f"(b[17]), "=f"(b[18]), "=f"(b[19]), "=f"(b[20]), "=f"(b[21]), "=f"(b[22]), "=f"(b[23]), "=f"(b[24]),
        "=f"(b[25]), "=f"(b[26]), "=f"(b[27]), "=f"(b[28]), "=f"(b[29]), "=f"(b[30]), "=f"(b[31])
      : "r"(taddr + 32u)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// two 16-column TMEM loads (this thread's lane), one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float (&a)[16], float (&b)[16]) {
  tmem_ld16_nowait(ta, a);
  tmem_ld16_nowait(tb, b);
  tmem_wait_ld();
}

// Thread layout: X_NP pipelines x X_EPI epilogue threads + 1 producer warp + X_ISSUERS MMA issuer warps (see fk_tc.cu for
// why dedicated issuer warps).  A lattice position (TMEM lane) is served by TWO epilogue threads -- thread ltid of the
// pipeline's first 128 and thread ltid of its second 128 (same TMEM lane quadrant, warp % 4) -- which take channels 0..15
// and 16..31: the hi/lo split makes this epilogue twice as long as the fp16 engine's, with two pipelines the tensor pipe sat
// idle 46 % of the time waiting for epilogues (profiles/r02_ncu_step_kernels.txt), and halving the per-thread instruction
// stream is what shortens them.  Barriers: fullA, fullB (weights landed), emptyA, emptyB (X_NP*X_EPI arrivals),
// mma[X_NP] (tcgen05.commit of the pipeline's current phase), ready[X_NP] (X_EPI arrivals: operand tiles written).
__global__ void __launch_bounds__(X_NP * X_EPI + 32 + X_ISSUERS * 32, 1) tcx_forward_kernel(TcxArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int pipe = tid / X_EPI, ltid = tid & 127, half = (tid >> 7) & 1;   // half: channels [16 half, 16 half + 16)
  const bool is_producer = warp == X_NP * (X_EPI / 32);
  const bool is_issuer = warp > X_NP * (X_EPI / 32);
  const int tile_bytes = 64 * a.npos;          // one fp16 tile: 4 channel groups x npos x 16 B
  const int slot_bytes = 2 * tile_bytes;       // hi tile, lo tile
  uint8_t* wA = smem;
  uint8_t* wB = smem + XA_BYTES;
  uint8_t* act0 = smem + XIMG_BYTES;
  uint8_t* tail = act0 + (size_t)X_NP * X_SLOTS * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);   // fullA, fullB, emptyA, emptyB, mma[2], ready[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 96);
  volatile uint32_t* issued = reinterpret_cast<volatile uint32_t*>(tail + 104);
  float* red = reinterpret_cast<float*>(tail + 128);    // [np][4 warps][2]
  TcBlockDesc* sdesc = reinterpret_cast<TcBlockDesc*>(tail + 384);
  uint8_t* ones_tile = smem + a.cst_off;
  uint8_t* zero_tile = ones_tile + 2048;

  const uint32_t fullA = smem_u32(&bars[0]), fullB = smem_u32(&bars[1]);
  const uint32_t emptyA = smem_u32(&bars[2]), emptyB = smem_u32(&bars[3]);
  const uint32_t mbar = smem_u32(&bars[4 + ((is_producer || is_issuer) ? 0 : pipe)]);
  const uint32_t rbar = smem_u32(&bars[6 + ((is_producer || is_issuer) ? 0 : pipe)]);

  if (tid == 32) {
    issued[0] = issued[1] = 0u;
    mbar_init(fullA, 1);
    mbar_init(fullB, 1);
    mbar_init(emptyA, X_NP * X_EPI);
    mbar_init(emptyB, X_NP * X_EPI);
    for (int p = 0; p < X_NP; ++p) {
      mbar_init(smem_u32(&bars[4 + p]), 1);
      mbar_init(smem_u32(&bars[6 + p]), X_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < a.nb; i += blockDim.x) sdesc[i] = a.desc[i];
  for (int i = tid; i < 256; i += blockDim.x)
    reinterpret_cast<uint4*>(ones_tile)[i] = i < 128 ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  {
    uint4* z = reinterpret_cast<uint4*>(act0);
    const int n16 = X_NP * X_SLOTS * slot_bytes / 16;
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool tile_mode = a.px.tiles != nullptr;
  const long long n_items = tile_mode ? *a.px.n_tiles : (a.wk.n_dev ? *a.wk.n_dev : a.n);
  const long long groups = (n_items + X_NP - 1) / X_NP;
  const long long my_iters = (long long)blockIdx.x < groups ? (groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (is_producer) {
    const long long total = my_iters * a.nb;
    for (long long s = 0; s < total; ++s) {
      if (lane == 0) {
        const uint8_t* src = a.images + (size_t)(s % a.nb) * XIMG_BYTES;
        if (s >= 1) mbar_wait(emptyA, (uint32_t)((s - 1) & 1));
        mbar_expect_tx(fullA, XA_BYTES);
        bulk_g2s(smem_u32(wA), src, XA_BYTES, fullA);
        if (s >= 1) mbar_wait(emptyB, (uint32_t)((s - 1) & 1));
        mbar_expect_tx(fullB, XB_BYTES);
        bulk_g2s(smem_u32(wB), src + XA_BYTES, XB_BYTES, fullB);
      }
      __syncwarp();
    }
  } else if (is_issuer) {
    const int role = __shfl_sync(0xffffffffu, warp - (X_NP * (X_EPI / 32) + 1), 0);
    const uint32_t tile16 = 4u * (uint32_t)a.npos, slot16 = 8u * (uint32_t)a.npos, kstep16 = 2u * (uint32_t)a.npos;
    const uint32_t hi = 8u | (1u << 14);                         // SBO = 8 (x16 B), descriptor version bit 46
    const uint32_t a_lbo = (uint32_t)a.npos << 16;
    const uint32_t act16 = smem_u32(act0) >> 4;
    const uint32_t wA16 = smem_u32(wA) >> 4, wB16 = smem_u32(wB) >> 4;
    const int P = a.P;
    // one tap, both k-steps: hi x [hi | lo] (N = 2n) then lo x hi (N = n) into the small-term columns
    auto tap_quad = [&](uint32_t d_tmem, uint32_t a16, int off, uint32_t w16, bool n32, uint32_t acc) {
      const uint32_t n = n32 ? 32u : 16u;
      const uint32_t ahi = ((a16 + (uint32_t)off) & 0x3FFFu) | a_lbo;
      const uint32_t alo = ((a16 + tile16 + (uint32_t)off) & 0x3FFFu) | a_lbo;
      const uint32_t blo = (w16 & 0x3FFFu) | ((2u * n) << 16);
      const uint32_t id2 = make_idesc((int)(2u * n)), id1 = make_idesc((int)n);
      const uint32_t wstep = n32 ? 128u : 64u;
      umma_f16_lohi(d_tmem, ahi, blo, hi, id2, acc);
      umma_f16_lohi(d_tmem + n, alo, blo, hi, id1, 1u);
      umma_f16_lohi(d_tmem, ahi + kstep16, blo + wstep, hi, id2, 1u);
      umma_f16_lohi(d_tmem + n, alo + kstep16, blo + wstep, hi, id1, 1u);
    };
    const uint32_t ones16 = smem_u32(ones_tile) >> 4, zero16 = smem_u32(zero_tile) >> 4;
    auto bias_mma = [&](uint32_t d_tmem, uint32_t bt16, int n) {
      const uint32_t alo = ones16 | ((zero16 - ones16) << 16);
      const uint32_t blo = (bt16 & 0x3FFFu) | (((zero16 - bt16) & 0x3FFFu) << 16);
      umma_f16_lohi(d_tmem, alo, blo, hi, make_idesc(n), 1u);
    };
    long long step = 0;
    uint32_t turn = 0, phase_count = 0;
    for (long long it = 0; it < my_iters; ++it) {
      for (int b = 0; b < a.nb; ++b, ++step) {
        const TcBlockDesc d = sdesc[b];
        bool have_a = false, have_b = false;
        const int last = __shfl_sync(0xffffffffu, d.last, 0);
        const int nph = last ? 4 : 3;
        for (int ph = 1; ph <= nph; ++ph, ++phase_count) {
          const uint32_t s0 = __shfl_sync(0xffffffffu, (uint32_t)(ph == 1 ? d.in_h : ph == 2 ? d.x1 : ph == 3 ? d.c : d.out_h), 0);
          const uint32_t s1 = __shfl_sync(0xffffffffu, (uint32_t)(ph == 1 ? d.in_v : d.out_a), 0);
          for (int p = 0; p < X_NP; ++p, ++turn) {
            if ((int)(turn % X_ISSUERS) != role) continue;
            if (ph == 1 && !have_a) { mbar_wait(fullA, (uint32_t)(step & 1)); have_a = true; }
            if (ph >= 2 && !have_b) { mbar_wait(fullB, (uint32_t)(step & 1)); have_b = true; }
            {
              uint32_t spins = 0;
              while (issued[p] < phase_count) {
                if (++spins > (1u << 26)) __trap();
              }
            }
            mbar_wait(smem_u32(&bars[6 + p]), phase_count & 1u);
            tc_fence_after();
            const uint32_t row00 = act16 + (uint32_t)(p * X_SLOTS) * slot16 + (uint32_t)a.p_first;
            const uint32_t dt = tmem_base + (uint32_t)(p * 128);
            const uint32_t a0 = row00 + s0 * slot16, a1 = row00 + s1 * slot16;
            if (elect_one()) {
              if (ph == 1) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  tap_quad(dt + 0, a0, j - 2, wA16 + XA_X / 16 + 256 * j, true, j != 0);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                  for (int j = 0; j < 3; ++j)
                    tap_quad(dt + 64, a1, (i - 2) * P + (j - 1), wA16 + XA_V / 16 + 256 * (i * 3 + j), true, (i | j) != 0);
                bias_mma(dt + 0, wA16 + XA_BT_X / 16, 32);
                bias_mma(dt + 64, wA16 + XA_BT_V / 16, 32);
              } else if (ph == 2) {
                tap_quad(dt + 0, a0, last ? -1 : 0, wB16 + XB_XX / 16, false, 0u);
                tap_quad(dt + 32, a1, -P, wB16 + XB_Y / 16, false, 0u);
                bias_mma(dt + 0, wB16 + XB_BT_XX / 16, 16);
                bias_mma(dt + 32, wB16 + XB_BT_Y / 16, 16);
              } else if (ph == 3) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                  for (int j = 0; j < 3; ++j)
                    tap_quad(dt + 0, a0, (i - 2) * P + (j - 2), wB16 + XB_H / 16 + 256 * (i * 3 + j), true, (i | j) != 0);
                bias_mma(dt + 0, wB16 + XB_BT_H / 16, 32);
              } else {
                tap_quad(dt + 0, a0, 0, wB16 + XB_HEAD / 16, false, 0u);
                bias_mma(dt + 0, wB16 + XB_BT_HEAD / 16, 16);
              }
              umma_commit(smem_u32(&bars[4 + p]));
              __threadfence_block();
              issued[p] = phase_count + 1u;
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // =============================== epilogue pipelines ===============================
    const uint32_t tm = tmem_base + (uint32_t)(pipe * 128) + ((uint32_t)((warp & 3) * 32) << 16);
    uint8_t* act = act0 + (size_t)pipe * X_SLOTS * slot_bytes;
    const int HW = a.H * a.W;
    const int bar_id = 1 + pipe;
    const int pos = a.p_first + ltid;
    const int prow = pos / a.P - 2, pcol = pos % a.P - 2;
    const int site = (pcol >= 0 && prow < a.H) ? prow * a.W + pcol : -1;
    float hres[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) hres[i] = 0.f;
    const uint32_t tmh = tm + (uint32_t)(16 * half);   // this thread's 16 columns inside a 32-column group

    // v (fp32, this thread's 16 channels) -> (hi, lo) fp16 rows of this thread's position in tile pair `slot`
    // `gd` (dump pass): the same (hi, lo) rows also go to the sample's cache tile pair at gd
    auto store_row = [&](int slot, const float* v, uint8_t* gd = nullptr) {
      uint8_t* bh = act + (size_t)slot * slot_bytes + (size_t)pos * 16 + (size_t)(2 * half) * a.npos * 16;
      uint8_t* bl = bh + tile_bytes;
      if (gd) gd += (size_t)pos * 16 + (size_t)(2 * half) * a.npos * 16;
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        uint4 qh, ql;
        float f[8];
        qh.x = pack_h2(v[8 * cg + 0], v[8 * cg + 1]);
        qh.y = pack_h2(v[8 * cg + 2], v[8 * cg + 3]);
        qh.z = pack_h2(v[8 * cg + 4], v[8 * cg + 5]);
        qh.w = pack_h2(v[8 * cg + 6], v[8 * cg + 7]);
        unpack_h8(qh, f);
        ql.x = pack_h2(v[8 * cg + 0] - f[0], v[8 * cg + 1] - f[1]);
        ql.y = pack_h2(v[8 * cg + 2] - f[2], v[8 * cg + 3] - f[3]);
        ql.z = pack_h2(v[8 * cg + 4] - f[4], v[8 * cg + 5] - f[5]);
        ql.w = pack_h2(v[8 * cg + 6] - f[6], v[8 * cg + 7] - f[7]);
        *reinterpret_cast<uint4*>(bh + (size_t)cg * a.npos * 16) = qh;
        *reinterpret_cast<uint4*>(bl + (size_t)cg * a.npos * 16) = ql;
        if (gd) {
          *reinterpret_cast<uint4*>(gd + (size_t)cg * a.npos * 16) = qh;
          *reinterpret_cast<uint4*>(gd + tile_bytes + (size_t)cg * a.npos * 16) = ql;
        }
      }
    };
    // Padding positions (site < 0) are never written: every slot was zeroed once at kernel start, an epilogue thread
    // only ever writes its own position, and the padding rows of every tensor are zero -- so they stay zero for the
    // whole kernel (the per-phase zero stores this replaces were 34 % of the LSU shared-memory wavefronts and all of the
    // store bank conflicts, profiles/r02_ncu_step_kernels.txt).
    auto load_row = [&](int slot, float* v) {
      const uint8_t* bh = act + (size_t)slot * slot_bytes + (size_t)pos * 16 + (size_t)(2 * half) * a.npos * 16;
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        float fh[8], fl[8];
        unpack_h8(*reinterpret_cast<const uint4*>(bh + (size_t)cg * a.npos * 16), fh);
        unpack_h8(*reinterpret_cast<const uint4*>(bh + tile_bytes + (size_t)cg * a.npos * 16), fl);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * cg + i] = fh[i] + fl[i];
      }
    };
    uint32_t mma_phase = 0;
    long long step = 0;

    if (tile_mode) {
      // =========================== tile pass of the prefix reuse (see TcxPrefix) ===========================
      const int W = a.W, H = a.H, P = a.P;
      const bool real_pos = pcol >= 0 && prow < a.px.rcap;          // a lattice column of a tile row inside the MMA range
      // idle positions (padding columns, rows past rcap) serve as loaders of segment A's halo, which lies above the MMA
      // range: idle thread number q < 2 W owns halo row q / W, column q % W
      int hq = -1;
      if (!real_pos) {
        int q = 0;
        for (int j = 0; j < ltid; ++j) {
          const int pj = a.p_first + j;
          const int rj = pj / P - 2, cj = pj % P - 2;
          q += (cj >= 0 && rj < a.px.rcap) ? 0 : 1;
        }
        if (q < 2 * W) hq = q;
      }
      const size_t half_off = (size_t)(2 * half) * a.npos * 16;
      // cached (hi, lo) rows of this thread's 16 channels at natural position lpos of tensor t of block b of sample smp
      auto halo_fetch = [&](int smp, int b, int t, int lpos, uint4 (&q)[4]) {
        const uint8_t* g = a.px.cache + (size_t)smp * a.px.cache_stride + (size_t)((b * 3 + t) * 2) * tile_bytes + half_off +
                           (size_t)lpos * 16;
        q[0] = __ldg(reinterpret_cast<const uint4*>(g));
        q[1] = __ldg(reinterpret_cast<const uint4*>(g + (size_t)a.npos * 16));
        q[2] = __ldg(reinterpret_cast<const uint4*>(g + tile_bytes));
        q[3] = __ldg(reinterpret_cast<const uint4*>(g + tile_bytes + (size_t)a.npos * 16));
      };
      auto halo_put = [&](int slot, int dpos, const uint4 (&q)[4]) {
        uint8_t* bh = act + (size_t)slot * slot_bytes + (size_t)dpos * 16 + half_off;
        *reinterpret_cast<uint4*>(bh) = q[0];
        *reinterpret_cast<uint4*>(bh + (size_t)a.npos * 16) = q[1];
        *reinterpret_cast<uint4*>(bh + tile_bytes) = q[2];
        *reinterpret_cast<uint4*>(bh + tile_bytes + (size_t)a.npos * 16) = q[3];
      };
      for (long long it = 0; it < my_iters; ++it) {
        const long long group = it * gridDim.x + blockIdx.x;
        const long long tix = group * X_NP + pipe;
        const bool active = tix < n_items;
        int cfgs[2] = {-1, -1}, smp[2] = {0, 0}, fa[2] = {-1, -1}, fb[2] = {-1, -1}, r0[2] = {0, 0}, kk[2] = {0, 0};
        if (active) {
          const int2 t = a.px.tiles[tix];
          cfgs[0] = t.x; cfgs[1] = t.y;
        }
#pragma unroll
        for (int sgi = 0; sgi < 2; ++sgi)
          if (cfgs[sgi] >= 0) {
            const TcWorkItem wi = a.wk.items[cfgs[sgi]];
            smp[sgi] = wi.sample;
            fa[sgi] = (int)wi.site_a;
            fb[sgi] = wi.site_b == 0xffffu ? -1 : (int)wi.site_b;
            const int ra = fa[sgi] / W, rb = fb[sgi] >= 0 ? fb[sgi] / W : ra;
            r0[sgi] = ra < rb ? ra : rb;
            kk[sgi] = H - r0[sgi];
          }
        // ---- role of this thread in this tile
        int seg = -1, lrow = 0;          // seg >= 0: this position is lattice site (lrow, pcol) of segment seg
        int h_smp = -1, h_row = 0, h_col = 0, h_dst = 0;   // halo duty: write the cached row of sample h_smp to position h_dst
        if (real_pos && cfgs[0] >= 0) {
          const int rowB0 = kk[0] + 2;
          if (prow < kk[0]) { seg = 0; lrow = prow + r0[0]; }
          else if (cfgs[1] >= 0) {
            if (prow < rowB0) { h_smp = smp[1]; h_row = r0[1] - 2 + (prow - kk[0]); h_col = pcol; h_dst = pos; }
            else if (prow < rowB0 + kk[1]) { seg = 1; lrow = prow - rowB0 + r0[1]; }
          }
        } else if (hq >= 0 && cfgs[0] >= 0) {
          h_smp = smp[0]; h_row = r0[0] - 2 + hq / W; h_col = hq % W; h_dst = (hq / W) * P + 2 + h_col;
        }
        const bool is_site = seg >= 0;
        const bool is_halo = h_smp >= 0;
        const bool halo_zero = is_halo && h_row < 0;               // above the lattice: the zero padding
        const int h_lpos = (h_row + 2) * P + h_col + 2;
        const int lsite = is_site ? lrow * W + pcol : -1;
        float sig = is_site ? (float)a.sigma[(size_t)smp[seg] * HW + lsite] : 0.f;
        if (is_site && (lsite == fa[seg] || lsite == fb[seg])) sig = -sig;
        {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
          if (half == 0) v[0] = sig;
          if (is_site) store_row(sdesc[0].in_v, v);
          if (is_halo) {   // the sample's own spins above r0 (exact in fp16: hi = +-1, lo = 0)
            const float hs = (!halo_zero && half == 0) ? (float)a.sigma[(size_t)h_smp * HW + h_row * W + h_col] : 0.f;
            uint4 q[4];
            q[0] = make_uint4(pack_h2(hs, 0.f), 0u, 0u, 0u);
            q[1] = q[2] = q[3] = make_uint4(0u, 0u, 0u, 0u);
            halo_put(sdesc[0].in_v, h_dst, q);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        for (int b = 0; b < a.nb; ++b, ++step) {
          const TcBlockDesc d = sdesc[b];
          // ================= phase 1 (halo rows: relu(v') and the residual sum come from the sample's cache)
          uint4 qa[4], qr[4];
          if (is_halo && !halo_zero) {
            halo_fetch(h_smp, b, 0, h_lpos, qa);
            if (d.out_r >= 0) halo_fetch(h_smp, b, 1, h_lpos, qr);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) qa[i] = qr[i] = make_uint4(0u, 0u, 0u, 0u);
          }
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          mbar_arrive(emptyA);
          if (is_halo) {
            if (d.out_r >= 0) halo_put(d.out_r, h_dst, qr);
            halo_put(d.out_a, h_dst, qa);
          }
          {
            float x[16], y[16];
            tmem_ld16x2(tmh + 0, tmh + 32, x, y);
            if (is_site) {
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + y[i]) * X_WINV, 0.f);
              store_row(d.x1, x);
            }
            tmem_ld16x2(tmh + 64, tmh + 96, x, y);
            if (is_site) {
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = (x[i] + y[i]) * X_WINV;
              if (d.out_r >= 0) {
                load_row(d.res_v, y);
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i] + x[i], 0.f);
                store_row(d.out_r, y);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], 0.f);
              store_row(d.out_a, x);
            }
          }
          // prefetch the concat halo while the phase-2 MMAs run
          if (is_halo && !halo_zero) halo_fetch(h_smp, b, 2, h_lpos, qa);
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(rbar);

          // ================= phase 2
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          if (is_halo) halo_put(d.c, h_dst, qa);
          {
            float x[16], y[16];
            tmem_ld16x2(tm + (uint32_t)(32 * half), tm + (uint32_t)(32 * half + 16), x, y);
            if (is_site) {
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + y[i]) * X_WINV, 0.f);
              store_row(d.c, x);
            }
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(rbar);

          // ================= phase 3
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          if (!d.last) mbar_arrive(emptyB);
          {
            float x[16], y[16];
            tmem_ld16x2(tmh + 0, tmh + 32, x, y);
            if (is_site) {
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = (x[i] + y[i]) * X_WINV;
              if (d.res_h >= 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] += hres[i];
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], 0.f);
              store_row(d.out_h, x);
              if (d.save_h) {
#pragma unroll
                for (int i = 0; i < 16; ++i) hres[i] = x[i];
              }
            }
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(rbar);

          // ================= phase 4 (last block): head, per-segment sums, local-energy terms
          if (d.last) {
            mbar_wait(mbar, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();
            mbar_arrive(emptyB);
            if (half == 0) {
              float sre = 0.f, sim = 0.f;
              {
                float v[32];
                tmem_ld32(tm + 0, v);
                if (is_site) {
                  const float re0 = (v[0] + v[16]) * X_WINV, re1 = (v[1] + v[17]) * X_WINV;
                  const float im0 = (v[2] + v[18]) * X_WINV, im1 = (v[3] + v[19]) * X_WINV;
                  const float x = 2.f * re0, y = 2.f * re1;
                  const float m = fmaxf(x, y);
                  const float half_lse = 0.5f * (m + logf(expf(x - m) + expf(y - m)));
                  const bool up = sig > 0.f;
                  const float2 base = *reinterpret_cast<const float2*>(a.px.rowcum + 2 * ((size_t)smp[seg] * HW + lsite));
                  sre = (up ? re0 : re1) - half_lse - base.x;      // difference to the sample's own term of this site
                  sim = (up ? im0 : im1) - base.y;
                }
              }
              float s_re[2], s_im[2];
#pragma unroll
              for (int sgi = 0; sgi < 2; ++sgi) {
                s_re[sgi] = seg == sgi ? sre : 0.f;
                s_im[sgi] = seg == sgi ? sim : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  s_re[sgi] += __shfl_xor_sync(0xffffffffu, s_re[sgi], o);
                  s_im[sgi] += __shfl_xor_sync(0xffffffffu, s_im[sgi], o);
                }
              }
              if (lane == 0) {
#pragma unroll
                for (int sgi = 0; sgi < 2; ++sgi) {
                  red[((pipe * 4 + (warp & 3)) * 2 + sgi) * 2 + 0] = s_re[sgi];
                  red[((pipe * 4 + (warp & 3)) * 2 + sgi) * 2 + 1] = s_im[sgi];
                }
              }
              tc_fence_before();
              named_sync(bar_id, 128);
              if (ltid < 2 && cfgs[ltid] >= 0) {   // thread 0 finishes segment A, thread 1 segment B
                const int sgi = ltid;
                float t0 = 0.f, t1 = 0.f;
                for (int w = 0; w < 4; ++w) {
                  t0 += red[((pipe * 4 + w) * 2 + sgi) * 2 + 0];
                  t1 += red[((pipe * 4 + w) * 2 + sgi) * 2 + 1];
                }
                const float dr = t0, di = t1;      // log psi(sigma') - log psi(sigma): the rows above r0 cancel exactly
                const float mag = expf(dr), m = a.wk.mel[cfgs[sgi]];
                float sn, cs;
                sincosf(di, &sn, &cs);
                atomicAdd(a.wk.eloc + 2 * smp[sgi], (double)m * (double)(mag * cs));
                atomicAdd(a.wk.eloc + 2 * smp[sgi] + 1, (double)m * (double)(mag * sn));
              }
              named_sync(bar_id, 128);   // `red` is reused by the next tile
            }
          }
        }
      }
    } else
    for (long long it = 0; it < my_iters; ++it) {
      const long long group = it * gridDim.x + blockIdx.x;
      const long long cfg = group * X_NP + pipe;
      const bool active = cfg < n_items;
      uint8_t* dump_cfg = (a.px.dump && active) ? a.px.dump + (size_t)cfg * a.px.cache_stride : nullptr;
      long long src = cfg;
      int flip_a = -1, flip_b = -1;
      if (a.wk.items && active) {   // connected configuration = the item's sample with the item's sites flipped
        const TcWorkItem wi = a.wk.items[cfg];
        src = wi.sample;
        flip_a = wi.site_a == 0xffffu ? -1 : (int)wi.site_a;
        flip_b = wi.site_b == 0xffffu ? -1 : (int)wi.site_b;
      }
      float sig = (active && site >= 0) ? (float)a.sigma[src * HW + site] : 0.f;
      if (site >= 0 && (site == flip_a || site == flip_b)) sig = -sig;
      {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        if (half == 0) v[0] = sig;
        store_row(sdesc[0].in_v, v);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(rbar);

      for (int b = 0; b < a.nb; ++b, ++step) {
        const TcBlockDesc d = sdesc[b];
        // ================= phase 1: 1x3 conv on h (cols 0..63) and 3x3 conv on v (cols 64..127)
        mbar_wait(mbar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        mbar_arrive(emptyA);   // chunk A is no longer read
        {
          float x[16], y[16];
          tmem_ld16x2(tmh + 0, tmh + 32, x, y);
          if (site >= 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + y[i]) * X_WINV, 0.f);
            store_row(d.x1, x);
          }
          tmem_ld16x2(tmh + 64, tmh + 96, x, y);
          if (site >= 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = (x[i] + y[i]) * X_WINV;
            if (d.out_r >= 0) {
              load_row(d.res_v, y);
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i] + x[i], 0.f);
              store_row(d.out_r, y, dump_cfg ? dump_cfg + (size_t)((b * 3 + 1) * 2) * tile_bytes : nullptr);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], 0.f);
            store_row(d.out_a, x, dump_cfg ? dump_cfg + (size_t)((b * 3 + 0) * 2) * tile_bytes : nullptr);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        // ================= phase 2: the two 1x1 convs -> concat (XX: cols 0..31, Y: cols 32..63; 16 large + 16 small each).
        // The first thread of a position takes XX = concat channels 0..15, the second Y = channels 16..31.
        mbar_wait(mbar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        {
          float x[16], y[16];
          tmem_ld16x2(tm + (uint32_t)(32 * half), tm + (uint32_t)(32 * half + 16), x, y);
          if (site >= 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaxf((x[i] + y[i]) * X_WINV, 0.f);
            store_row(d.c, x, dump_cfg ? dump_cfg + (size_t)((b * 3 + 2) * 2) * tile_bytes : nullptr);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        // ================= phase 3: 3x3 conv on the concat tensor (cols 0..63), residual, relu
        mbar_wait(mbar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if (!d.last) mbar_arrive(emptyB);
        {
          float x[16], y[16];
          tmem_ld16x2(tmh + 0, tmh + 32, x, y);
          if (site >= 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = (x[i] + y[i]) * X_WINV;
            if (d.res_h >= 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] += hres[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], 0.f);
            store_row(d.out_h, x);
            if (d.save_h) {
#pragma unroll
              for (int i = 0; i < 16; ++i) hres[i] = x[i];
            }
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        // ================= phase 4 (last block): head 1x1 conv (C -> 4) + normalisation + combine
        if (d.last) {
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          mbar_arrive(emptyB);
          if (half == 0) {   // 4 real logit columns: the first thread of every position finishes the configuration
          float sre = 0.f, sim = 0.f;
          {
            float v[32];
            tmem_ld32(tm + 0, v);
            if (site >= 0) {
              const float re0 = (v[0] + v[16]) * X_WINV, re1 = (v[1] + v[17]) * X_WINV;
              const float im0 = (v[2] + v[18]) * X_WINV, im1 = (v[3] + v[19]) * X_WINV;
              const float x = 2.f * re0, y = 2.f * re1;
              const float m = fmaxf(x, y);
              const float half_lse = 0.5f * (m + logf(expf(x - m) + expf(y - m)));
              const bool up = sig > 0.f;
              sre = (up ? re0 : re1) - half_lse;
              sim = up ? im0 : im1;
              if (a.px.siteterm && active)
                *reinterpret_cast<float2*>(a.px.siteterm + 2 * ((size_t)cfg * HW + site)) = make_float2(sre, sim);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sre += __shfl_xor_sync(0xffffffffu, sre, o);
            sim += __shfl_xor_sync(0xffffffffu, sim, o);
          }
          if (lane == 0) {
            red[(pipe * 4 + (warp & 3)) * 2 + 0] = sre;
            red[(pipe * 4 + (warp & 3)) * 2 + 1] = sim;
          }
          tc_fence_before();
          named_sync(bar_id, 128);
          if (ltid == 0 && active) {
            float r0 = 0.f, r1 = 0.f;
            for (int w = 0; w < 4; ++w) {
              r0 += red[(pipe * 4 + w) * 2 + 0];
              r1 += red[(pipe * 4 + w) * 2 + 1];
            }
            if (a.wk.eloc) {   // fused local-energy term: ratio in complex64, accumulation in complex128
              const float dr = r0 - a.wk.logpsi0[2 * src], di = r1 - a.wk.logpsi0[2 * src + 1];
              const float mag = expf(dr), m = a.wk.mel[cfg];
              float sn, cs;
              sincosf(di, &sn, &cs);
              atomicAdd(a.wk.eloc + 2 * src, (double)m * (double)(mag * cs));
              atomicAdd(a.wk.eloc + 2 * src + 1, (double)m * (double)(mag * sn));
            } else {
              a.out[2 * cfg + 0] = r0;
              a.out[2 * cfg + 1] = r1;
            }
          }
          named_sync(bar_id, 128);   // `red` is reused by the next configuration
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct TcxGeometry { int P, p_first, npos; size_t smem_bytes, tail; bool ok; };

static TcxGeometry tcx_geometry(const fk_net* net) {
  TcxGeometry g;
  g.P = net->W + 2;
  g.p_first = 2 * g.P + 2;
  const int p_last = (net->H + 1) * g.P + net->W + 1;
  const int T = (p_last - g.p_first + 1 + 127) / 128;
  g.npos = ((g.p_first + 128 + 2) + 7) / 8 * 8;
  g.tail = (384 + sizeof(TcBlockDesc) * (size_t)(2 * net->depth - 2) + 64 + 127) / 128 * 128;
  g.smem_bytes = (size_t)XIMG_BYTES + (size_t)X_NP * X_SLOTS * 128 * g.npos + g.tail + 4096;
  g.ok = T == 1 && g.smem_bytes <= 227 * 1024;
  return g;
}

int tcx_supported(const fk_net* net) {
  if (net->kind != FK_NET_CONV2D || net->C != 32 || net->k != 3) return 0;
  return tcx_geometry(net).ok ? 1 : 0;
}

static size_t x256(size_t x) { return (x + 255) / 256 * 256; }
static size_t tcx_desc_offset(int nb) { return x256((size_t)nb * XIMG_BYTES); }
static size_t tcx_pack_offset(int nb) { return tcx_desc_offset(nb) + x256(sizeof(TcBlockDesc) * nb); }

// allocation + the wiring tables (called once, from fk_net_create)
int tcx_prepare(fk_net* net) {
  if (!tcx_supported(net)) return 0;
  const int nb = 2 * net->depth - 2;
  FK_CHECK_CUDA(cudaMalloc(&net->d_tc_exact, tcx_pack_offset(nb) + sizeof(XPackDesc) * nb));
  // slot wiring with the horizontal residual in registers (3 slots): in-place rules of fk_tc.cu::tc_pack_weights
  std::vector<TcBlockDesc> desc(nb);
  int rc[X_SLOTS] = {0, 0, 0};
  auto get = [&]() {
    for (int i = 0; i < X_SLOTS; ++i)
      if (rc[i] == 0) { rc[i] = 1; return i; }
    return -1;
  };
  const int in = get();
  rc[in]++;
  int v = in, h = in, v_pair = -1;
  for (int b = 0; b < nb; ++b) {
    const bool last = (b == nb - 1);
    const bool res2 = (b >= 2 && b % 2 == 0 && !last);
    if (b % 2 == 1 && !last) { v_pair = v; rc[v]++; }
    TcBlockDesc d;
    d.in_v = (int8_t)v; d.in_h = (int8_t)h; d.last = last ? 1 : 0; d.pad1 = 0;
    d.save_h = (int8_t)(((b + 1) % 2 == 1 && b + 1 != nb - 1) ? 1 : 0);
    int x1;
    if (rc[h] == 1) { x1 = h; } else { x1 = get(); rc[h]--; }
    int a1;
    if (rc[v] == 1) { a1 = v; } else { a1 = get(); rc[v]--; }
    FK_REQUIRE(x1 >= 0 && a1 >= 0, "tcx_prepare: out of shared-memory activation slots");
    d.out_r = d.res_v = (int8_t)(res2 ? v_pair : -1);
    const int c = x1;
    int v_next = a1;
    d.res_h = -1;
    if (res2) { rc[a1]--; v_next = v_pair; d.res_h = 0; }
    d.x1 = (int8_t)x1; d.out_a = (int8_t)a1; d.c = (int8_t)c; d.out_h = (int8_t)c;
    desc[b] = d;
    v = v_next; h = c;
  }
  FK_CHECK_CUDA(cudaMemcpy((uint8_t*)net->d_tc_exact + tcx_desc_offset(nb), desc.data(), sizeof(TcBlockDesc) * nb, cudaMemcpyHostToDevice));
  std::vector<XPackDesc> pd(nb);
  for (int b = 0; b < nb; ++b) {
    const ConvOp* o = &net->ops[5 * b];   // program order per block: V, X, XX, Y, H
    for (int r = 0; r < 5; ++r) { pd[b].w[r] = o[r].w_off; pd[b].b[r] = o[r].b_off; }
    pd[b].w_head = net->ops.back().w_off; pd[b].b_head = net->ops.back().b_off;
    pd[b].cin = b == 0 ? 1 : 32;
  }
  FK_CHECK_CUDA(cudaMemcpy((uint8_t*)net->d_tc_exact + tcx_pack_offset(nb), pd.data(), sizeof(XPackDesc) * nb, cudaMemcpyHostToDevice));
  return 0;
}

// ---- prefix reuse: work list -> tiles ------------------------------------------------------------------------------------
constexpr int XP_MAXK = 64;    // row classes (k = rows to recompute = H - r0)
constexpr int XP_MAXRUN = 2 * XP_MAXK + 2;
struct TcxPlan {
  int count[XP_MAXK + 1];      // configurations per class
  int off[XP_MAXK + 2];        // class k occupies list[off[k] .. off[k + 1])
  int cursor[XP_MAXK + 1];     // scatter cursors
  int nruns;
  int run_a[XP_MAXRUN], run_sa[XP_MAXRUN], run_b[XP_MAXRUN], run_sb[XP_MAXRUN], run_tile0[XP_MAXRUN + 1];
  long long n_tiles;
};

__device__ __forceinline__ int xp_rows(const TcWorkItem& wi, int W, int H) {
  const int ra = (int)wi.site_a / W, rb = wi.site_b == 0xffffu ? ra : (int)wi.site_b / W;
  return H - (ra < rb ? ra : rb);
}

__global__ void tcx_plan_zero_kernel(TcxPlan* plan) {
  for (int i = threadIdx.x; i <= XP_MAXK; i += blockDim.x) { plan->count[i] = 0; plan->cursor[i] = 0; }
}

__global__ void tcx_class_count_kernel(const TcWorkItem* __restrict__ items, const long long* __restrict__ n_dev, int W, int H,
                                       TcxPlan* plan) {
  __shared__ int hist[XP_MAXK + 1];
  for (int i = threadIdx.x; i <= XP_MAXK; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const long long n = *n_dev;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (long long)gridDim.x * blockDim.x)
    atomicAdd(&hist[xp_rows(items[f], W, H)], 1);
  __syncthreads();
  for (int i = threadIdx.x; i <= XP_MAXK; i += blockDim.x)
    if (hist[i]) atomicAdd(&plan->count[i], hist[i]);
}

// one thread: class offsets and the pairing of the classes.  Two segments of kA and kB rows share a tile when
// kA + 2 + kB <= rcap (two halo rows between them); greedy: the smallest non-empty class takes partners from the largest class
// that still fits, classes that fit with nobody become single-segment tiles.
__global__ void tcx_plan_kernel(TcxPlan* plan, int H, int rcap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int rem[XP_MAXK + 1], cur[XP_MAXK + 1];
  int o = 0;
  for (int k = 0; k <= XP_MAXK; ++k) {
    plan->off[k] = o;
    o += plan->count[k];
    rem[k] = plan->count[k];
    cur[k] = 0;
  }
  plan->off[XP_MAXK + 1] = o;
  const int T = rcap - 2;
  int nr = 0, tile = 0;
  auto add = [&](int ca, int sa, int cb, int sb, int m) {
    if (m <= 0) return;
    plan->run_a[nr] = ca; plan->run_sa[nr] = sa; plan->run_b[nr] = cb; plan->run_sb[nr] = sb; plan->run_tile0[nr] = tile;
    ++nr;
    tile += m;
  };
  int lo = 1, hi = H < XP_MAXK ? H : XP_MAXK;
  for (;;) {
    while (lo <= hi && rem[lo] == 0) ++lo;
    while (hi >= lo && rem[hi] == 0) --hi;
    if (lo > hi) break;
    if (lo == hi) {
      if (2 * lo <= T) {
        const int m = rem[lo] / 2;
        add(lo, cur[lo], lo, cur[lo] + m, m);
        cur[lo] += 2 * m; rem[lo] -= 2 * m;
      }
      add(lo, cur[lo], -1, 0, rem[lo]);
      rem[lo] = 0;
      break;
    }
    if (lo + hi <= T) {
      const int m = rem[lo] < rem[hi] ? rem[lo] : rem[hi];
      add(lo, cur[lo], hi, cur[hi], m);
      cur[lo] += m; rem[lo] -= m; cur[hi] += m; rem[hi] -= m;
    } else {
      add(hi, cur[hi], -1, 0, rem[hi]);
      cur[hi] += rem[hi]; rem[hi] = 0;
    }
  }
  plan->run_tile0[nr] = tile;
  plan->nruns = nr;
  plan->n_tiles = tile;
}

__global__ void tcx_class_scatter_kernel(const TcWorkItem* __restrict__ items, const long long* __restrict__ n_dev, int W, int H,
                                         TcxPlan* plan, int* __restrict__ list) {
  const long long n = *n_dev;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (long long)gridDim.x * blockDim.x) {
    const int k = xp_rows(items[f], W, H);
    list[plan->off[k] + atomicAdd(&plan->cursor[k], 1)] = (int)f;
  }
}

__global__ void tcx_tiles_kernel(const TcxPlan* __restrict__ plan, const int* __restrict__ list, int2* __restrict__ tiles) {
  const long long n = plan->n_tiles;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    int r = 0;
    while (r + 1 < plan->nruns && plan->run_tile0[r + 1] <= t) ++r;
    const int i = (int)(t - plan->run_tile0[r]);
    const int ca = plan->run_a[r], cb = plan->run_b[r];
    int2 out;
    out.x = list[plan->off[ca] + plan->run_sa[r] + i];
    out.y = cb >= 0 ? list[plan->off[cb] + plan->run_sb[r] + i] : -1;
    tiles[t] = out;
  }
}

int tcx_pack_weights(fk_net* net, cudaStream_t s) {
  if (!net->d_tc_exact) return 0;
  const int nb = 2 * net->depth - 2;
  const XPackDesc* d_pd = reinterpret_cast<const XPackDesc*>((uint8_t*)net->d_tc_exact + tcx_pack_offset(nb));
  tcx_pack_kernel<<<nb, 256, 0, s>>>(net->d_weff, d_pd, (uint8_t*)net->d_tc_exact);
  FK_CHECK_LAUNCH();
  return 0;
}

static int tcx_launch(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, cudaStream_t s, const TcWork* work,
                      const TcxPrefix* px);

int tcx_log_psi(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, cudaStream_t s, const TcWork* work) {
  return tcx_launch(net, sigma, n, log_psi_out, s, work, nullptr);
}

static int tcx_launch(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, cudaStream_t s, const TcWork* work,
                      const TcxPrefix* px) {
  FK_REQUIRE(net->params_set && net->d_tc_exact, "tc-exact engine: not available for this machine, or parameters never set");
  if (n == 0) return 0;
  const TcxGeometry g = tcx_geometry(net);
  FK_REQUIRE(g.ok, "tc-exact engine: lattice %dx%d does not fit one 128-row tile", net->H, net->W);
  const int nb = 2 * net->depth - 2;
  TcxArgs a;
  a.images = (const uint8_t*)net->d_tc_exact;
  a.desc = reinterpret_cast<const TcBlockDesc*>((const uint8_t*)net->d_tc_exact + tcx_desc_offset(nb));
  a.sigma = sigma; a.out = log_psi_out; a.n = n;
  a.H = net->H; a.W = net->W; a.P = g.P; a.nb = nb; a.npos = g.npos; a.p_first = g.p_first;
  a.cst_off = (int)(g.smem_bytes - 4096);
  if (work) a.wk = *work; else a.wk = TcWork{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (px) a.px = *px; else a.px = TcxPrefix{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
  int dev = 0, sms = 148;
  FK_CHECK_CUDA(cudaGetDevice(&dev));
  FK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long groups = (n + X_NP - 1) / X_NP;
  const unsigned grid = (unsigned)std::min<long long>(groups, sms);
  FK_CHECK_CUDA(cudaFuncSetAttribute(tcx_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  tcx_forward_kernel<<<grid, X_NP * X_EPI + 32 + X_ISSUERS * 32, g.smem_bytes, s>>>(a);
  FK_CHECK_LAUNCH();
  return 0;
}

// ---- local energy with prefix reuse: dump pass over the samples, tiles from the work list, tile pass -------------------------
int xp_rcap(const fk_net* net) { return (128 - net->W) / (net->W + 2) + 1; }

int xp_geometry_ok(const fk_net* net) {
  const int rcap = xp_rcap(net);
  if (net->H > XP_MAXK || net->H > rcap) return 0;
  if (128 - rcap * net->W < 2 * net->W) return 0;      // not enough idle positions to load segment A's halo
  return 1;
}

int tcx_prefix_supported(const fk_net* net) { return tcx_supported(net) && xp_geometry_ok(net) ? 1 : 0; }

// workspace: plan | class lists (cap ints) | tiles (cap int2)
int64_t xp_tiles_workspace_bytes(int64_t cap) { return (int64_t)(x256(sizeof(TcxPlan)) + x256((size_t)cap * 4) + x256((size_t)cap * 8)); }

int xp_build_tiles(const fk_net* net, const TcWork* work, int64_t cap, void* ws, const int2** tiles_out, const long long** n_tiles_out,
                   cudaStream_t s) {
  uint8_t* base = (uint8_t*)ws;
  TcxPlan* plan = reinterpret_cast<TcxPlan*>(base);
  int* list = reinterpret_cast<int*>(base + x256(sizeof(TcxPlan)));
  int2* tiles = reinterpret_cast<int2*>(base + x256(sizeof(TcxPlan)) + x256((size_t)cap * 4));
  tcx_plan_zero_kernel<<<1, 128, 0, s>>>(plan);
  FK_CHECK_LAUNCH();
  tcx_class_count_kernel<<<296, 256, 0, s>>>(work->items, work->n_dev, net->W, net->H, plan);
  FK_CHECK_LAUNCH();
  tcx_plan_kernel<<<1, 32, 0, s>>>(plan, net->H, xp_rcap(net));
  FK_CHECK_LAUNCH();
  tcx_class_scatter_kernel<<<296, 256, 0, s>>>(work->items, work->n_dev, net->W, net->H, plan, list);
  FK_CHECK_LAUNCH();
  tcx_tiles_kernel<<<296, 256, 0, s>>>(plan, list, tiles);
  FK_CHECK_LAUNCH();
  *tiles_out = tiles;
  *n_tiles_out = &plan->n_tiles;
  return 0;
}

struct XpLayout { size_t cache, siteterm, tiles_ws, total, stride; };
static XpLayout xp_layout(const fk_net* net, int64_t B, int64_t cap) {
  const TcxGeometry g = tcx_geometry(net);
  const int nb = 2 * net->depth - 2;
  XpLayout L;
  L.stride = (size_t)nb * 3 * 2 * 64 * g.npos;
  size_t o = 0;
  L.cache = o; o = x256(o + (size_t)B * L.stride);
  L.siteterm = o; o = x256(o + (size_t)B * net->sites * 8);
  L.tiles_ws = o; o = x256(o + (size_t)xp_tiles_workspace_bytes(cap));
  L.total = o;
  return L;
}

int64_t tcx_prefix_workspace_bytes(const fk_net* net, int64_t B, int64_t cap) { return (int64_t)xp_layout(net, B, cap).total; }

// `work` holds the device work list of the B samples (items, count, matrix elements), eloc accumulators initialised with the
// diagonal terms; log psi of the samples is written to work->logpsi0 by the dump pass
int tcx_local_energy_prefix(fk_net* net, const int8_t* sigma, int64_t B, int64_t cap, const TcWork* work, void* ws, int64_t ws_bytes,
                            cudaStream_t s) {
  FK_REQUIRE(tcx_prefix_supported(net), "tc-exact prefix reuse: machine outside the envelope");
  const XpLayout L = xp_layout(net, B, cap);
  FK_REQUIRE((int64_t)L.total <= ws_bytes, "tc-exact prefix reuse: workspace too small (%lld < %zu)", (long long)ws_bytes, L.total);
  uint8_t* base = (uint8_t*)ws;
  float* siteterm = reinterpret_cast<float*>(base + L.siteterm);
  const int rcap = xp_rcap(net);
  // dump pass: log psi of the samples + their activation cache + per-site terms
  TcxPrefix pd = {nullptr, nullptr, nullptr, nullptr, base + L.cache, siteterm, (long long)L.stride, rcap};
  if (tcx_launch(net, sigma, B, const_cast<float*>(work->logpsi0), s, nullptr, &pd)) return 1;
  const int2* tiles = nullptr;
  const long long* n_tiles = nullptr;
  if (xp_build_tiles(net, work, cap, base + L.tiles_ws, &tiles, &n_tiles, s)) return 1;
  // tile pass
  TcxPrefix pt = {tiles, n_tiles, base + L.cache, siteterm, nullptr, nullptr, (long long)L.stride, rcap};
  return tcx_launch(net, sigma, cap, nullptr, s, work, &pt);
}

}  // namespace fk
