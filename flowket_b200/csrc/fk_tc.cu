// Tensor-core engine: the whole ConvNetAutoregressive2D forward (all 2*depth-2 blocks + head + log-space
// normalisation + one-hot combine) of a configuration in ONE persistent kernel on tcgen05 tensor cores.
//
//   * activations never leave the SM: fp16 tiles in shared memory in the UMMA canonical K-major / no-swizzle
//     layout [channel group of 8][padded position][8 channels]; a conv tap is a *shifted view* of the same
//     tile (descriptor start address + tap offset * 16 B), so the 3x3 / 1x3 / shifted 1x1 masked convolutions
//     are implicit GEMMs with M = 128 padded positions, N = 32 (or 16), K = 16 per tcgen05.mma, no im2col;
//   * accumulators live in TMEM; 128 epilogue threads (one TMEM lane = one lattice position each) apply
//     bias / residual / relu / zero-padding mask, round to fp16 and write the next layer's operand tile;
//   * weights: one pre-packed fp16 image per block (UMMA canonical layout), double-buffered in shared memory
//     and streamed from L2 with cp.async.bulk + mbarrier complete_tx while the previous block computes;
//   * up to three independent pipelines (128 threads each) per CTA process one configuration each; while one
//     pipeline's epilogue runs, the others' MMAs occupy the tensor core; a producer warp refills the weight ring
//     through full/empty mbarriers, so the pipelines never synchronise with each other;
//   * the last epilogue fuses the head: logits -> log-space normalisation -> select by sigma -> sum over sites.
//
// Semantics: machines/conv_net_autoregressive_2D.py:24-74, machines/abstract_machine.py:31-57,
// deepar/layers/autoregressive.py:7-22 (same wiring as the fp32 layer program in fk_net.cu).
// Numerics: fp16 operands (11-bit significand: 8x tighter than bf16 at the same tensor-core rate; activations are
// O(1) post-relu values, conversions saturate), fp32 accumulation / bias / residual / normalisation.
// Tolerance stated in DESIGN.md and tests/test_gpu_tc.py.
#include <algorithm>

#include "fk_net.cuh"
#include "fk_tc_common.cuh"

namespace fk {

constexpr int TC_SLOTS = 4;

struct TcBlockDesc {
  int8_t in_v, in_h, out_a, out_r, res_v, x1, c, out_h, res_h, last, save_h, pad1;
};

struct TcPackDesc {
  long long w[5], b[5];  // V, X, XX, Y, H offsets into the effective-weight buffer
  long long w_head, b_head;
  int cin;               // 1 for the first block
};

// ---- weight packing: fp32 effective weights [tap][ci][co] -> fp16 UMMA canonical K-major B tiles -----------
__global__ void tc_pack_kernel(const float* __restrict__ weff, const TcPackDesc* __restrict__ pd,
                               uint8_t* __restrict__ images) {
  const TcPackDesc d = pd[blockIdx.x];
  uint8_t* img = images + (size_t)blockIdx.x * IMG_BYTES;
  __half* w16 = reinterpret_cast<__half*>(img);
  // region table: (byte offset, taps, N, op index, cin)
  const int reg_off[6] = {IMG_V, IMG_X, IMG_XX, IMG_Y, IMG_H, IMG_HEAD};
  const int reg_taps[6] = {9, 3, 1, 1, 9, 1};
  const int reg_n[6] = {32, 32, 16, 16, 32, 16};
  for (int r = 0; r < 6; ++r) {
    const int N = reg_n[r], taps = reg_taps[r];
    const int cin = (r == 0 || r == 1) ? d.cin : 32;
    const int nreal = (r == 5) ? 4 : N;
    const long long woff = (r == 5) ? d.w_head : d.w[r];
    const int total = taps * 2 * 2 * N * 8;  // (tap, kstep, kgroup, n, e)
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int el = e & 7;
      const int n = (e >> 3) % N;
      const int g = ((e >> 3) / N) & 1;
      const int ks = ((e >> 3) / N / 2) & 1;
      const int tap = (e >> 3) / N / 4;
      const int ci = ks * 16 + g * 8 + el;
      float v = 0.f;
      if (ci < cin && n < nreal) v = weff[woff + ((long long)tap * cin + ci) * nreal + n];
      w16[reg_off[r] / 2 + e] = __float2half_rn(v);
    }
  }
  float* bias = reinterpret_cast<float*>(img + IMG_BIAS);
  const int boff[5] = {0, 32, 64, 80, 96}, bn[5] = {32, 32, 16, 16, 32};
  for (int r = 0; r < 5; ++r)
    for (int i = threadIdx.x; i < bn[r]; i += blockDim.x) bias[boff[r] + i] = weff[d.b[r] + i];
  float* hb = reinterpret_cast<float*>(img + IMG_HEAD_BIAS);
  for (int i = threadIdx.x; i < 16; i += blockDim.x) hb[i] = i < 4 ? weff[d.b_head + i] : 0.f;
  // bias tiles (B operands of the bias MMAs): per output channel [hi, lo, 0 x 6] with hi + lo = bias to ~2^-22
  for (int i = threadIdx.x; i < 144; i += blockDim.x) {
    float bv;
    int off;   // byte offset of the channel's 16-byte row
    if (i < 32) { bv = weff[d.b[1] + i]; off = IMG_BT_XV + 16 * i; }                           // X -> columns 0..31
    else if (i < 64) { bv = weff[d.b[0] + i - 32]; off = IMG_BT_XV + 16 * i; }                 // V -> columns 32..63
    else if (i < 80) { bv = weff[d.b[2] + i - 64]; off = IMG_BT_XXY + 16 * (i - 64); }         // XX
    else if (i < 96) { bv = weff[d.b[3] + i - 80]; off = IMG_BT_XXY + 16 * (i - 64); }         // Y
    else if (i < 128) { bv = weff[d.b[4] + i - 96]; off = IMG_BT_H + 16 * (i - 96); }          // H
    else { bv = (i - 128) < 4 ? weff[d.b_head + i - 128] : 0.f; off = IMG_BT_HEAD + 16 * (i - 128); }
    const __half hi = __float2half_rn(bv);
    const __half lo = __float2half_rn(bv - __half2float(hi));
    __half* row = reinterpret_cast<__half*>(img + off);
    row[0] = hi; row[1] = lo;
    for (int k = 2; k < 8; ++k) row[k] = __float2half_rn(0.f);
  }
}

struct TcArgs {
  const uint8_t* images;
  const TcBlockDesc* desc;
  const int8_t* sigma;
  float* out;
  long long n;
  int H, W, P, nb, T, npos, p_first, np, tmem_cols, cst_off, slots;
  // activation dump for the tensor-core gradient (T == 1 only): fp16 tiles [cfg][nb*5 + 1][64*npos] in the shared-memory
  // tile layout (tensor order per block: x1, relu(v'), residual v, concat, h_out; last tile = input), relu masks
  // [cfg][nb][5][128] (bit c = channel c active, 0 on padding rows) and logits [cfg][128][4]
  uint8_t* dump;
  uint32_t* dump_mask;
  float* dump_logits;
  TcWork wk;   // local-energy work list (all null: plain log psi of the configurations in `sigma`)
  TcxPrefix px; // prefix reuse, tile pass (fk_net.cuh); the cache is the gradient dump of the samples (tiles relu(v'), residual v,
                // concat = tensors 1, 2, 3 of every block); all null = off
};

constexpr int TC_DUMP_TENSORS = 5;
constexpr int TC_MAX_T = 3;
constexpr int TC_MAX_NP = 3;

#ifdef FK_TC_TRACE
__device__ long long fk_tc_trace_buf[3 * 40 * 16];
#define TRACE(slot) do { if (!DUMP && ltid == 0 && blockIdx.x == 0 && it == 3) fk_tc_trace_buf[(pipe * 40 + b) * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do {} while (0)
#endif

// Thread layout: np pipelines x 128 epilogue threads (one configuration each) + 1 producer warp (weight images)
//                + TC_ISSUERS issuer warps (tcgen05.mma).
// Why issuer warps (measured: tools/umma_bench*.cu and the clock64 trace of this kernel, tests/tools_tc_trace.py):
// tcgen05.mma takes its descriptors from uniform registers.  Issued under `if (thread == 0)` inside an epilogue warp
// the compiler wraps every MMA in a per-thread R2UR loop (>100 cycles per MMA, and the pipelines fall into lockstep:
// all epilogues run together while the tensor pipe idles); a converged warp with one elected lane issues back-to-back,
// and with two or more issuers in flight the pipe retires an N=32 MMA every 40 cycles -- the shared-memory operand
// bandwidth (128 B/cycle).  The (block, phase, pipeline) turns are dealt round-robin to the issuers.
// Barriers: full[2] (weights landed, tx-count), empty[2] (np*128 arrivals: every epilogue thread is done with the
//           image), mma[np] (tcgen05.commit of the pipeline's current phase), ready[np] (128 arrivals: the operand
//           tiles of the pipeline's next phase are written and fenced).
constexpr int TC_ISSUERS = 3;   // measured on the 10x10 lattice: 2 -> 2.47, 3 -> 2.59, 4 -> 2.38 M configurations/s
// HREG (two-tile lattices, e.g. 12x12): the horizontal stack's residual-pair input -- each epilogue thread's own rows,
// 32 fp16 per tile -- lives in registers instead of a fourth shared-memory tile, so that two pipelines fit where one
// did.  MAXNP sizes the launch bound (registers per thread): 3 pipelines -> 128 registers, 2 -> 168.
template <bool DUMP, bool HREG, int MAXNP>
__global__ void __launch_bounds__(MAXNP * 128 + 32 + TC_ISSUERS * 32, 1) tc_forward_kernel(TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int pipe = tid >> 7, ltid = tid & 127;
  const bool is_producer = warp == a.np * 4;
  const bool is_issuer = warp > a.np * 4;
  const int buf_bytes = 64 * a.npos;  // 4 channel groups x npos x 16 B
  uint8_t* wbuf = smem;
  uint8_t* act0 = smem + 2 * IMG_BYTES;
  uint8_t* tail = act0 + (size_t)a.np * a.slots * buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);           // full[0..1], empty[2..3], mma[4..6], ready[7..9]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 96);
  volatile uint32_t* issued = reinterpret_cast<volatile uint32_t*>(tail + 104);   // [np] phases issued per pipeline
  float* red = reinterpret_cast<float*>(tail + 128);             // [np][4 warps][2]
  int* taps = reinterpret_cast<int*>(tail + 256);                // tap offsets in positions: V[9] H[9] X[3]
  TcBlockDesc* sdesc = reinterpret_cast<TcBlockDesc*>(tail + 384);
  // A operand of the bias MMAs: 128 rows of [1, 1, 0 x 6] (first K half) + 2 KB of zeros (second K half of A and of the
  // bias tiles).  Last region of the dynamic shared memory, so every LBO that points at the zeros is positive.
  uint8_t* ones_tile = smem + a.cst_off;
  uint8_t* zero_tile = ones_tile + 2048;

  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]);
  const uint32_t mbar = smem_u32(&bars[4 + ((is_producer || is_issuer) ? 0 : pipe)]);
  const uint32_t rbar = smem_u32(&bars[7 + ((is_producer || is_issuer) ? 0 : pipe)]);

  if (tid == 32) {  // (not warp 0: it must reach the .sync.aligned TMEM allocation converged)
    issued[0] = issued[1] = issued[2] = 0u;
    mbar_init(full0, 1);
    mbar_init(full0 + 8, 1);
    mbar_init(empty0, a.np * 128);
    mbar_init(empty0 + 8, a.np * 128);
    for (int p = 0; p < a.np; ++p) {
      mbar_init(smem_u32(&bars[4 + p]), 1);
      mbar_init(smem_u32(&bars[7 + p]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        taps[i * 3 + j] = (i - 2) * a.P + (j - 1);      // 3x3 on the vertical stack: pad top 2, left 1, right 1
        taps[9 + i * 3 + j] = (i - 2) * a.P + (j - 2);  // 3x3 on the concat tensor: pad top 2, left 2
      }
    for (int j = 0; j < 3; ++j) taps[18 + j] = j - 2;   // 1x3 on the horizontal stack: pad left 2
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < a.nb; i += blockDim.x) sdesc[i] = a.desc[i];
  for (int i = tid; i < 256; i += blockDim.x) {
    reinterpret_cast<uint4*>(ones_tile)[i] = i < 128 ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);   // fp16 1.0, 1.0
  }
  {  // zero every activation tile once: padding rows / slack positions are never written afterwards
    uint4* z = reinterpret_cast<uint4*>(act0);
    const int n16 = a.np * a.slots * buf_bytes / 16;
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool tile_mode = !DUMP && a.px.tiles != nullptr;    // (T == 1 only, checked by the launcher)
  const long long n_items = tile_mode ? *a.px.n_tiles : (a.wk.n_dev ? *a.wk.n_dev : a.n);
  const long long groups = (n_items + a.np - 1) / a.np;  // configuration groups (one per CTA iteration)
  const long long my_iters = (long long)blockIdx.x < groups ? (groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (is_producer) {
    // =============================== weight producer ===============================
    const long long total = my_iters * a.nb;
    for (long long s = 0; s < total; ++s) {   // whole warp walks the loop (it must reach the final barrier converged)
      if (lane == 0) {
        const uint32_t sel = (uint32_t)(s & 1);
        if (s >= 2) mbar_wait(empty0 + 8 * sel, (uint32_t)(((s >> 1) - 1) & 1));
        mbar_expect_tx(full0 + 8 * sel, IMG_BYTES);
        bulk_g2s(smem_u32(wbuf + (size_t)sel * IMG_BYTES), a.images + (size_t)(s % a.nb) * IMG_BYTES, IMG_BYTES,
                 full0 + 8 * sel);
      }
      __syncwarp();
    }
  } else if (is_issuer) {
    // =============================== MMA issuers (whole warp walks the loops, one elected lane issues) ==========
    // The (block, phase, pipeline) turns are dealt round-robin to the issuers.  Issuers run unordered with respect to
    // each other except per pipeline: phase k+1 of a pipeline is looked at only after its phase k has been issued
    // (issued[p] counter) -- otherwise the parity test of the ready barrier could alias two phases.
    // Everything that feeds a descriptor goes through __shfl_sync(.., 0) once per turn so that the compiler keeps
    // the per-tap arithmetic in uniform registers.
    const int role = __shfl_sync(0xffffffffu, warp - (a.np * 4 + 1), 0);
    const uint32_t buf16 = 4u * (uint32_t)a.npos, kstep16 = 2u * (uint32_t)a.npos;   // 16-byte units
    const uint32_t idesc32 = make_idesc(32), idesc16 = make_idesc(16);
    const uint32_t hi = 8u | (1u << 14);                         // SBO = 8 (x16 B), descriptor version bit 46
    const uint32_t a_lbo = (uint32_t)a.npos << 16;               // LBO of the activation tiles
    const uint32_t act16 = smem_u32(act0) >> 4;
    const int P = a.P;
    // one (tap, both k-steps) pair: A = view of the tile shifted by `off` positions, B = weight tiles of the tap
    auto tap_pair = [&](uint32_t d_tmem, uint32_t a16, int off, uint32_t w16, bool n32, uint32_t acc) {
      const uint32_t alo = ((a16 + (uint32_t)off) & 0x3FFFu) | a_lbo;
      const uint32_t blo = (w16 & 0x3FFFu) | ((n32 ? 32u : 16u) << 16);
      umma_f16_lohi(d_tmem, alo, blo, hi, n32 ? idesc32 : idesc16, acc);
      umma_f16_lohi(d_tmem, alo + kstep16, blo + (n32 ? 64u : 32u), hi, n32 ? idesc32 : idesc16, 1u);
    };
    // bias MMA: D[:, col0 .. col0+N) += ones x bias tile (one K=16 step; the second K halves are the zero tile)
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
        const uint32_t wsel = (uint32_t)(step & 1);
        bool have_w = false;
        const uint32_t wimg16 = __shfl_sync(0xffffffffu, smem_u32(wbuf + (size_t)wsel * IMG_BYTES) >> 4, 0);
        const int last = __shfl_sync(0xffffffffu, d.last, 0);
        const int nph = last ? 4 : 3;
        for (int ph = 1; ph <= nph; ++ph, ++phase_count) {
          // operand tiles of this phase
          const uint32_t s0 = __shfl_sync(0xffffffffu, (uint32_t)(ph == 1 ? d.in_h : ph == 2 ? d.x1 : ph == 3 ? d.c : d.out_h), 0);
          const uint32_t s1 = __shfl_sync(0xffffffffu, (uint32_t)(ph == 1 ? d.in_v : d.out_a), 0);
          for (int p = 0; p < a.np; ++p, ++turn) {
            if ((int)(turn % TC_ISSUERS) != role) continue;
            if (!have_w) {
              mbar_wait(full0 + 8 * wsel, (uint32_t)((step >> 1) & 1));
              have_w = true;
            }
            if (TC_ISSUERS > 1) {
              uint32_t spins = 0;
              while (issued[p] < phase_count) {
                if (++spins > (1u << 26)) __trap();
              }
            }
            mbar_wait(smem_u32(&bars[7 + p]), phase_count & 1u);
            tc_fence_after();
            const uint32_t row00 = act16 + (uint32_t)(p * a.slots) * buf16 + (uint32_t)a.p_first;
            const bool leader = elect_one();
            for (int t = 0; t < a.T; ++t) {
              const uint32_t dt = tmem_base + (uint32_t)((p * a.T + t) * 128);
              const uint32_t a0 = row00 + (uint32_t)(t * 128) + s0 * buf16, a1 = row00 + (uint32_t)(t * 128) + s1 * buf16;
              if (leader) {
                if (ph == 1) {
#pragma unroll
                  for (int j = 0; j < 3; ++j)      // 1x3 on the horizontal stack: pad left 2
                    tap_pair(dt + 0, a0, j - 2, wimg16 + IMG_X / 16 + 128 * j, true, j != 0);
#pragma unroll
                  for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)    // 3x3 on the vertical stack: pad top 2, left 1, right 1
                      tap_pair(dt + 32, a1, (i - 2) * P + (j - 1), wimg16 + IMG_V / 16 + 128 * (i * 3 + j), true, (i | j) != 0);
                  bias_mma(dt + 0, wimg16 + IMG_BT_XV / 16, 64);
                } else if (ph == 2) {
                  tap_pair(dt + 64, a0, last ? -1 : 0, wimg16 + IMG_XX / 16, false, 0u);
                  tap_pair(dt + 80, a1, -P, wimg16 + IMG_Y / 16, false, 0u);
                  bias_mma(dt + 64, wimg16 + IMG_BT_XXY / 16, 32);
                } else if (ph == 3) {
#pragma unroll
                  for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)    // 3x3 on the concat tensor: pad top 2, left 2
                      tap_pair(dt + 96, a0, (i - 2) * P + (j - 2), wimg16 + IMG_H / 16 + 128 * (i * 3 + j), true, (i | j) != 0);
                  bias_mma(dt + 96, wimg16 + IMG_BT_H / 16, 32);
                } else {
                  tap_pair(dt + 0, a0, 0, wimg16 + IMG_HEAD / 16, false, 0u);
                  bias_mma(dt + 0, wimg16 + IMG_BT_HEAD / 16, 16);
                }
              }
            }
            if (leader) {
              umma_commit(smem_u32(&bars[4 + p]));
              if (TC_ISSUERS > 1) {
                __threadfence_block();
                issued[p] = phase_count + 1u;
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // =============================== epilogue pipelines ===============================
    const uint32_t tm_pipe = tmem_base + (uint32_t)(pipe * a.T * 128);
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* act = act0 + (size_t)pipe * a.slots * buf_bytes;
    const uint32_t act16 = smem_u32(act) >> 4;      // everything below is in 16-byte units
    const uint32_t buf16 = 4u * (uint32_t)a.npos;
    const uint32_t kstep16 = 2u * (uint32_t)a.npos;  // two channel groups per k-step
    const int HW = a.H * a.W;
    const uint32_t idesc32 = make_idesc(32), idesc16 = make_idesc(16);
    const uint64_t adesc0 = make_desc(0, a.npos, 8);
    const uint64_t bdesc32 = make_desc(0, 32, 8), bdesc16 = make_desc(0, 16, 8);
    const int bar_id = 1 + pipe;

    int pos[TC_MAX_T], site[TC_MAX_T];
#pragma unroll
    for (int t = 0; t < TC_MAX_T; ++t) {
      pos[t] = a.p_first + t * 128 + ltid;
      const int r = pos[t] / a.P - 2, c = pos[t] % a.P - 2;
      site[t] = (t < a.T && c >= 0 && r < a.H) ? r * a.W + c : -1;
    }

    constexpr int HT = HREG ? 2 : 1;
    uint32_t hres[HT][16];   // HREG: packed fp16 rows (one per tile) of the horizontal residual-pair input
#pragma unroll
    for (int t = 0; t < HT; ++t)
#pragma unroll
      for (int i = 0; i < 16; ++i) hres[t][i] = 0u;

    auto store_row = [&](int slot, int p, const float* v) {
      uint8_t* base = act + (size_t)slot * buf_bytes + (size_t)p * 16;
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        uint4 q;
        q.x = pack_h2(v[8 * cg + 0], v[8 * cg + 1]);
        q.y = pack_h2(v[8 * cg + 2], v[8 * cg + 3]);
        q.z = pack_h2(v[8 * cg + 4], v[8 * cg + 5]);
        q.w = pack_h2(v[8 * cg + 6], v[8 * cg + 7]);
        *reinterpret_cast<uint4*>(base + (size_t)cg * a.npos * 16) = q;
      }
    };
    auto load_row = [&](int slot, int p, float* v) {
      const uint8_t* base = act + (size_t)slot * buf_bytes + (size_t)p * 16;
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        const uint4 q = *reinterpret_cast<const uint4*>(base + (size_t)cg * a.npos * 16);
        unpack_h8(q, v + 8 * cg);
      }
    };
    // gradient dump: this thread's row of tile `tensor` of block `b` (fp16, tile layout) + its relu mask
    auto dump_row = [&](long long cfg, int b, int tensor, const float* v, bool valid) {
      const size_t tile_bytes = (size_t)64 * a.npos;
      const int tile_idx = b < 0 ? a.nb * TC_DUMP_TENSORS : b * TC_DUMP_TENSORS + tensor;
      uint8_t* tile = a.dump + ((size_t)cfg * (a.nb * TC_DUMP_TENSORS + 1) + tile_idx) * tile_bytes;
      uint32_t mask = 0;
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        uint4 q;
        q.x = pack_h2(v[8 * cg + 0], v[8 * cg + 1]);
        q.y = pack_h2(v[8 * cg + 2], v[8 * cg + 3]);
        q.z = pack_h2(v[8 * cg + 4], v[8 * cg + 5]);
        q.w = pack_h2(v[8 * cg + 6], v[8 * cg + 7]);
        *reinterpret_cast<uint4*>(tile + ((size_t)cg * a.npos + pos[0]) * 16) = q;
        // rows outside the computed range are never written by an epilogue: keep them zero for the dW gathers
        if (ltid < a.p_first) *reinterpret_cast<uint4*>(tile + ((size_t)cg * a.npos + ltid) * 16) = make_uint4(0, 0, 0, 0);
        if (ltid < a.npos - a.p_first - 128)
          *reinterpret_cast<uint4*>(tile + ((size_t)cg * a.npos + a.p_first + 128 + ltid) * 16) = make_uint4(0, 0, 0, 0);
      }
      if (b >= 0) {
#pragma unroll
        for (int c = 0; c < 32; ++c) mask |= (valid && v[c] > 0.f) ? (1u << c) : 0u;
        a.dump_mask[(((size_t)cfg * a.nb + b) * TC_DUMP_TENSORS + tensor) * 128 + ltid] = mask;
      }
    };
    uint32_t mma_phase = 0;
    long long step = 0;  // weight-pipeline step: block (step % nb) lives in buffer (step & 1)

    if (tile_mode) {
      // =========================== tile pass of the prefix reuse (TcxPrefix, fk_net.cuh; one M tile) ===========================
      const int W = a.W, H = a.H, P = a.P;
      const int p0 = pos[0];
      const int prow = p0 / P - 2, pcol = p0 % P - 2;
      const bool real_pos = pcol >= 0 && prow < a.px.rcap;
      int hq = -1;            // idle positions load segment A's halo (two rows above the MMA range): halo row hq / W, column hq % W
      if (!real_pos) {
        int q = 0;
        for (int j = 0; j < ltid; ++j) {
          const int pj = a.p_first + j;
          q += (pj % P - 2 >= 0 && pj / P - 2 < a.px.rcap) ? 0 : 1;
        }
        if (q < 2 * W) hq = q;
      }
      const size_t tile_b = (size_t)64 * a.npos;
      auto halo_fetch = [&](int smp, int b, int tensor, int lpos, uint4 (&q)[4]) {
        const uint8_t* g = a.px.cache + (size_t)smp * a.px.cache_stride + (size_t)(b * TC_DUMP_TENSORS + tensor) * tile_b + (size_t)lpos * 16;
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) q[cg] = __ldg(reinterpret_cast<const uint4*>(g + (size_t)cg * a.npos * 16));
      };
      auto put_q = [&](int slot, int dpos, const uint4 (&q)[4]) {
        uint8_t* base = act + (size_t)slot * buf_bytes + (size_t)dpos * 16;
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) *reinterpret_cast<uint4*>(base + (size_t)cg * a.npos * 16) = q[cg];
      };
      for (long long it = 0; it < my_iters; ++it) {
        const long long group = it * gridDim.x + blockIdx.x;
        const long long tix = group * a.np + pipe;
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
        int seg = -1, lrow = 0;
        int h_smp = -1, h_row = 0, h_col = 0, h_dst = 0;
        if (real_pos && cfgs[0] >= 0) {
          const int rowB0 = kk[0] + 2;
          if (prow < kk[0]) { seg = 0; lrow = prow + r0[0]; }
          else if (cfgs[1] >= 0) {
            if (prow < rowB0) { h_smp = smp[1]; h_row = r0[1] - 2 + (prow - kk[0]); h_col = pcol; h_dst = p0; }
            else if (prow < rowB0 + kk[1]) { seg = 1; lrow = prow - rowB0 + r0[1]; }
          }
        } else if (hq >= 0 && cfgs[0] >= 0) {
          h_smp = smp[0]; h_row = r0[0] - 2 + hq / W; h_col = hq % W; h_dst = (hq / W) * P + 2 + h_col;
        }
        const bool is_site = seg >= 0;
        const bool is_halo = h_smp >= 0;
        const bool halo_zero = is_halo && h_row < 0;
        const bool own_halo = is_halo && h_dst == p0;               // in-tile halo row (segment B): replaces this position's own store
        const int h_lpos = (h_row + 2) * P + h_col + 2;
        const int lsite = is_site ? lrow * W + pcol : -1;
        float sg1 = is_site ? (float)a.sigma[(size_t)smp[seg] * HW + lsite] : 0.f;
        if (is_site && (lsite == fa[seg] || lsite == fb[seg])) sg1 = -sg1;
        uint4 qa[4], qr[4];
        {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
          v[0] = sg1;
          if (!own_halo) store_row(sdesc[0].in_v, p0, v);
          if (is_halo) {
            const float hs = halo_zero ? 0.f : (float)a.sigma[(size_t)h_smp * HW + h_row * W + h_col];
            qa[0] = make_uint4(pack_h2(hs, 0.f), 0u, 0u, 0u);
            qa[1] = qa[2] = qa[3] = make_uint4(0u, 0u, 0u, 0u);
            put_q(sdesc[0].in_v, h_dst, qa);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        for (int b = 0; b < a.nb; ++b, ++step) {
          const TcBlockDesc d = sdesc[b];
          const uint32_t wsel = (uint32_t)(step & 1);
          mbar_wait(full0 + 8 * wsel, (uint32_t)((step >> 1) & 1));
          // ================= phase 1
          if (is_halo && !halo_zero) {
            halo_fetch(h_smp, b, 1, h_lpos, qa);
            if (d.out_r >= 0) halo_fetch(h_smp, b, 2, h_lpos, qr);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) qa[i] = qr[i] = make_uint4(0u, 0u, 0u, 0u);
          }
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          if (is_halo) {
            if (d.out_r >= 0) put_q(d.out_r, h_dst, qr);
            put_q(d.out_a, h_dst, qa);
          }
          {
            float v[32];
            tmem_ld32(tm_pipe + lane_sel + 0, v);
            if (is_site) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            store_row(d.x1, p0, v);
            tmem_ld32(tm_pipe + lane_sel + 32, v);
            if (is_site) {
              if (d.out_r >= 0) {
                float r[32];
                load_row(d.res_v, p0, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = fmaxf(r[i] + v[i], 0.f);
                store_row(d.out_r, p0, r);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              store_row(d.out_a, p0, v);
            } else if (!own_halo) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
              if (d.out_r >= 0) store_row(d.out_r, p0, v);
              store_row(d.out_a, p0, v);
            }
          }
          if (is_halo && !halo_zero) halo_fetch(h_smp, b, 3, h_lpos, qa);   // concat halo, prefetched during the phase-2 MMAs
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(rbar);

          // ================= phase 2
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          if (is_halo) put_q(d.c, h_dst, qa);
          {
            float v[32];
            tmem_ld32(tm_pipe + lane_sel + 64, v);
            if (is_site) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (!own_halo) store_row(d.c, p0, v);
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(rbar);

          // ================= phase 3
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          {
            float v[32];
            tmem_ld32(tm_pipe + lane_sel + 96, v);
            if (is_site) {
              if (d.res_h >= 0) {
                float r[32];
                load_row(d.res_h, p0, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += r[i];
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            store_row(d.out_h, p0, v);
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(rbar);

          // ================= phase 4 (last block): head, per-segment sums, local-energy terms
          if (d.last) {
            mbar_wait(mbar, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();
            float sre = 0.f, sim = 0.f;
            {
              float v[16];
              tmem_ld16(tm_pipe + lane_sel + 0, v);
              if (is_site) {
                const float re0 = v[0], re1 = v[1], im0 = v[2], im1 = v[3];
                const float x = 2.f * re0, y = 2.f * re1;
                const float m = fmaxf(x, y);
                const float half_lse = 0.5f * (m + logf(expf(x - m) + expf(y - m)));
                const bool up = sg1 > 0.f;
                const float2 base = *reinterpret_cast<const float2*>(a.px.rowcum + 2 * ((size_t)smp[seg] * HW + lsite));
                sre = (up ? re0 : re1) - half_lse - base.x;      // difference to the sample's own term of this site
                sim = (up ? im0 : im1) - base.y;
              }
            }
#pragma unroll
            for (int sgi = 0; sgi < 2; ++sgi) {
              float tr = seg == sgi ? sre : 0.f, ti = seg == sgi ? sim : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                tr += __shfl_xor_sync(0xffffffffu, tr, o);
                ti += __shfl_xor_sync(0xffffffffu, ti, o);
              }
              if (lane == 0) {
                red[(pipe * 4 + (warp & 3)) * 2 + 0] = tr;
                red[(pipe * 4 + (warp & 3)) * 2 + 1] = ti;
              }
              tc_fence_before();
              named_sync(bar_id, 128);
              if (ltid == 0 && cfgs[sgi] >= 0) {
                float t0 = 0.f, t1 = 0.f;
                for (int w = 0; w < 4; ++w) {
                  t0 += red[(pipe * 4 + w) * 2 + 0];
                  t1 += red[(pipe * 4 + w) * 2 + 1];
                }
                const float dr = t0, di = t1;      // log psi(sigma') - log psi(sigma): the rows above r0 cancel exactly
                const float mag = expf(dr), m = a.wk.mel[cfgs[sgi]];
                float sn, cs;
                sincosf(di, &sn, &cs);
                atomicAdd(a.wk.eloc + 2 * smp[sgi], (double)m * (double)(mag * cs));
                atomicAdd(a.wk.eloc + 2 * smp[sgi] + 1, (double)m * (double)(mag * sn));
              }
              named_sync(bar_id, 128);   // `red` is reused by the next segment / tile
            }
          }
          mbar_arrive(empty0 + 8 * wsel);
        }
      }
    } else
    for (long long it = 0; it < my_iters; ++it) {
      const long long group = it * gridDim.x + blockIdx.x;
      const long long cfg = group * a.np + pipe;
      const bool active = cfg < n_items;

      // ---- input embedding: channel 0 = sigma, channels 1..31 = 0, padding positions = 0
      // (work list: the configuration is the item's sample with the item's sites flipped -- an exchange of an
      //  anti-parallel pair flips both, operators/heisenberg.py:94-95)
      float sig[TC_MAX_T];
      long long src = cfg;
      int flip_a = -1, flip_b = -1;
      if (a.wk.items && active) {
        const TcWorkItem wi = a.wk.items[cfg];
        src = wi.sample;
        flip_a = wi.site_a == 0xffffu ? -1 : (int)wi.site_a;
        flip_b = wi.site_b == 0xffffu ? -1 : (int)wi.site_b;
      }
      {
        const int in_slot = sdesc[0].in_v;
#pragma unroll
        for (int t = 0; t < TC_MAX_T; ++t) {
          if (t >= a.T) break;
          sig[t] = (active && site[t] >= 0) ? (float)a.sigma[src * HW + site[t]] : 0.f;
          if (site[t] >= 0 && (site[t] == flip_a || site[t] == flip_b)) sig[t] = -sig[t];
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
          v[0] = sig[t];
          store_row(in_slot, pos[t], v);
          if (DUMP && active && t == 0) dump_row(cfg, -1, 0, v, site[t] >= 0);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(rbar);   // operand tile of block 0 is ready

      for (int b = 0; b < a.nb; ++b, ++step) {
        const TcBlockDesc d = sdesc[b];
        const uint32_t wsel = (uint32_t)(step & 1);
        mbar_wait(full0 + 8 * wsel, (uint32_t)((step >> 1) & 1));
        const uint8_t* wimg = wbuf + (size_t)wsel * IMG_BYTES;
        const uint32_t wimg16 = smem_u32(wimg) >> 4;

        // ================= phase 1: 1x3 conv on h (cols 0..31) and 3x3 conv on v (cols 32..63)
        TRACE(0);
        mbar_wait(mbar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        TRACE(2);
#pragma unroll
        for (int t = 0; t < TC_MAX_T; ++t) {
          if (t >= a.T) break;
          float v[32];
          tmem_ld32(tm_pipe + lane_sel + (uint32_t)(t * 128) + 0, v);
          if (site[t] >= 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          store_row(d.x1, pos[t], v);
          if (DUMP && active && t == 0) dump_row(cfg, b, 0, v, site[t] >= 0);
          tmem_ld32(tm_pipe + lane_sel + (uint32_t)(t * 128) + 32, v);
          if (site[t] >= 0) {
            if (d.out_r >= 0) {
              float r[32];
              load_row(d.res_v, pos[t], r);
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = fmaxf(r[i] + v[i], 0.f);
              store_row(d.out_r, pos[t], r);
              if (DUMP && active && t == 0) dump_row(cfg, b, 2, r, true);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
            if (d.out_r >= 0) {
              store_row(d.out_r, pos[t], v);
              if (DUMP && active && t == 0) dump_row(cfg, b, 2, v, false);
            }
          }
          store_row(d.out_a, pos[t], v);
          if (DUMP && active && t == 0) dump_row(cfg, b, 1, v, site[t] >= 0);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        // ================= phase 2: the two 1x1 convs (C -> C/2): x1 (RightShift in the last block) and
        //                   DownShift(relu(v')) -> concat tensor (cols 64..95)
        TRACE(3);
        mbar_wait(mbar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        TRACE(5);
#pragma unroll
        for (int t = 0; t < TC_MAX_T; ++t) {
          if (t >= a.T) break;
          float v[32];
          tmem_ld32(tm_pipe + lane_sel + (uint32_t)(t * 128) + 64, v);
          if (site[t] >= 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          store_row(d.c, pos[t], v);
          if (DUMP && active && t == 0) dump_row(cfg, b, 3, v, site[t] >= 0);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        // ================= phase 3: 3x3 conv on the concat tensor (cols 96..127), residual, relu
        TRACE(6);
        mbar_wait(mbar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        TRACE(8);
#pragma unroll
        for (int t = 0; t < TC_MAX_T; ++t) {
          if (t >= a.T) break;
          float v[32];
          tmem_ld32(tm_pipe + lane_sel + (uint32_t)(t * 128) + 96, v);
          if (site[t] >= 0) {
            if (d.res_h >= 0) {
              float r[32];
              if (HREG) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hres[t < HT ? t : 0][i]));
                  r[2 * i] = f.x;
                  r[2 * i + 1] = f.y;
                }
              } else {
                load_row(d.res_h, pos[t], r);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += r[i];
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          store_row(d.out_h, pos[t], v);
          if (HREG && d.save_h && t < HT) {   // this block's output is the next pair's input: keep the stored fp16 values
#pragma unroll
            for (int i = 0; i < 16; ++i) hres[t < HT ? t : 0][i] = pack_h2(v[2 * i], v[2 * i + 1]);
          }
          if (DUMP && active && t == 0) dump_row(cfg, b, 4, v, site[t] >= 0);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(rbar);

        TRACE(9);
        // ================= phase 4 (last block): head 1x1 conv (C -> 4) + normalisation + combine
        if (d.last) {
          mbar_wait(mbar, mma_phase);
          mma_phase ^= 1;
          tc_fence_after();
          float sre = 0.f, sim = 0.f;
#pragma unroll
          for (int t = 0; t < TC_MAX_T; ++t) {
            if (t >= a.T) break;
            float v[16];
            tmem_ld16(tm_pipe + lane_sel + (uint32_t)(t * 128) + 0, v);
            if (site[t] >= 0) {
              const float re0 = v[0], re1 = v[1], im0 = v[2], im1 = v[3];   // (head bias added by the bias MMA)
              if (DUMP && active && t == 0)
                *reinterpret_cast<float4*>(a.dump_logits + ((size_t)cfg * 128 + ltid) * 4) = make_float4(re0, re1, im0, im1);
              const float x = 2.f * re0, y = 2.f * re1;
              const float m = fmaxf(x, y);
              const float half_lse = 0.5f * (m + logf(expf(x - m) + expf(y - m)));
              const bool up = sig[t] > 0.f;  // class 0 <-> sigma = +1
              sre += (up ? re0 : re1) - half_lse;
              sim += up ? im0 : im1;
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
        }
        // this thread is done with the weight image (the block's MMAs have retired, its bias reads are behind it)
        mbar_arrive(empty0 + 8 * wsel);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct TcGeometry {
  int P, p_first, T, npos, np, tmem_cols, slots;
  bool hreg;
  size_t smem_bytes;
  bool ok;
};

static TcGeometry tc_geometry(const fk_net* net) {
  TcGeometry g;
  g.P = net->W + 2;
  g.p_first = 2 * g.P + 2;
  const int p_last = (net->H + 1) * g.P + net->W + 1;  // position of site (H-1, W-1)
  g.T = (p_last - g.p_first + 1 + 127) / 128;
  g.npos = ((g.p_first + g.T * 128 + 2) + 7) / 8 * 8;
  const size_t buf = (size_t)64 * g.npos;
  const size_t tail = (384 + sizeof(TcBlockDesc) * (size_t)(2 * net->depth - 2) + 64 + 127) / 128 * 128;
  g.ok = g.T <= TC_MAX_T;
  auto fit = [&](int slots, size_t* smem_out) {   // pipelines that fit with `slots` activation tiles each
    for (int np = TC_MAX_NP; np >= 1; --np) {
      const size_t smem = 2 * (size_t)IMG_BYTES + (size_t)np * slots * buf + tail + 4096;   // + ones / zero tiles
      if (smem <= 227 * 1024 && np * g.T * 128 <= 512) { *smem_out = smem; return np; }
    }
    *smem_out = 0;
    return 0;
  };
  size_t smem4 = 0, smem3 = 0;
  const int np4 = fit(TC_SLOTS, &smem4);
  const int np3 = g.T == 2 ? std::min(2, fit(TC_SLOTS - 1, &smem3)) : 0;   // register residual: two-tile lattices only
  g.hreg = np3 > np4;
  g.slots = g.hreg ? TC_SLOTS - 1 : TC_SLOTS;
  g.np = g.hreg ? np3 : np4;
  if (g.hreg) smem3 = 2 * (size_t)IMG_BYTES + (size_t)g.np * g.slots * buf + tail + 4096;
  g.smem_bytes = g.hreg ? smem3 : smem4;
  if (g.np == 0) { g.ok = false; g.np = 1; }
  int cols = g.np * g.T * 128;
  g.tmem_cols = cols <= 128 ? 128 : (cols <= 256 ? 256 : 512);
  return g;
}

int tc_supported(const fk_net* net) {
  if (net->kind != FK_NET_CONV2D || net->C != 32 || net->k != 3) return 0;
  return tc_geometry(net).ok ? 1 : 0;
}

// device buffer layout: [nb weight images][nb TcBlockDesc][nb TcPackDesc], each region 256 B aligned
static size_t align256z(size_t x) { return (x + 255) / 256 * 256; }
static size_t tc_desc_offset(int nb) { return align256z((size_t)nb * IMG_BYTES); }
static size_t tc_pack_offset(int nb) { return tc_desc_offset(nb) + align256z(sizeof(TcBlockDesc) * nb); }

int tc_prepare(fk_net* net) {
  if (!tc_supported(net)) return 0;
  const int nb = 2 * net->depth - 2;
  const size_t total = tc_pack_offset(nb) + sizeof(TcPackDesc) * nb;
  {
    FK_CHECK_CUDA(cudaMalloc(&net->d_tc_weights, total));
    net->tc_weight_bytes = (int64_t)total;
    // residual / buffer wiring -> shared-memory slots (reference counting; residual adds are done in place)
    std::vector<TcBlockDesc> desc(nb);
    // In-place rules (safe because an epilogue thread only touches its own position and every MMA of the
    // phase has retired before the epilogue starts): x1 may overwrite h, relu(v') may overwrite v, the concat
    // tensor overwrites x1, h' overwrites the concat tensor (or the pair input when it carries the residual).
    // With g.hreg the horizontal pair input is kept in registers by the epilogue threads (res_h is then only a flag
    // and save_h marks the blocks whose output is a pair input), which frees one tile.
    const TcGeometry g = tc_geometry(net);
    int rc[TC_SLOTS] = {0, 0, 0, 0};
    auto get = [&]() {
      for (int i = 0; i < g.slots; ++i)
        if (rc[i] == 0) { rc[i] = 1; return i; }
      return -1;
    };
    const int in = get();
    rc[in]++;
    int v = in, h = in, v_pair = -1, h_pair = -1;
    for (int b = 0; b < nb; ++b) {
      const bool last = (b == nb - 1);
      const bool res2 = (b >= 2 && b % 2 == 0 && !last);
      if (b % 2 == 1 && !last) {
        v_pair = v; rc[v]++;
        if (!g.hreg) { h_pair = h; rc[h]++; }
      }
      TcBlockDesc d;
      d.in_v = (int8_t)v; d.in_h = (int8_t)h; d.last = last ? 1 : 0; d.pad1 = 0;
      d.save_h = (int8_t)((g.hreg && (b + 1) % 2 == 1 && b + 1 != nb - 1) ? 1 : 0);   // block b+1 pins its input
      int x1;
      if (rc[h] == 1) { x1 = h; } else { x1 = get(); rc[h]--; }
      int a1;
      if (rc[v] == 1) { a1 = v; } else { a1 = get(); rc[v]--; }
      FK_REQUIRE(x1 >= 0 && a1 >= 0, "tc_pack_weights: out of shared-memory activation slots");
      d.out_r = d.res_v = (int8_t)(res2 ? v_pair : -1);
      const int c = x1;
      int v_next = a1, hn = c;
      d.res_h = -1;
      if (res2) {
        rc[a1]--; v_next = v_pair;            // relu(v') only feeds the 1x1 conv; the residual sum replaces the pair input
        if (g.hreg) { d.res_h = 0; }
        else { rc[c]--; hn = h_pair; d.res_h = (int8_t)h_pair; }
      }
      d.x1 = (int8_t)x1; d.out_a = (int8_t)a1; d.c = (int8_t)c; d.out_h = (int8_t)hn;
      desc[b] = d;
      v = v_next; h = hn;
    }
    FK_CHECK_CUDA(cudaMemcpy((uint8_t*)net->d_tc_weights + tc_desc_offset(nb), desc.data(), sizeof(TcBlockDesc) * nb,
                             cudaMemcpyHostToDevice));
    std::vector<TcPackDesc> pd(nb);
    for (int b = 0; b < nb; ++b) {
      const ConvOp* o = &net->ops[5 * b];
      // program order per block: V, X, XX, Y, H
      for (int r = 0; r < 5; ++r) { pd[b].w[r] = o[r].w_off; pd[b].b[r] = o[r].b_off; }
      pd[b].w_head = net->ops.back().w_off; pd[b].b_head = net->ops.back().b_off;
      pd[b].cin = b == 0 ? 1 : 32;
    }
    FK_CHECK_CUDA(cudaMemcpy((uint8_t*)net->d_tc_weights + tc_pack_offset(nb), pd.data(), sizeof(TcPackDesc) * nb,
                             cudaMemcpyHostToDevice));
  }
  return 0;
}

int tc_pack_weights(fk_net* net, cudaStream_t s) {
  const int nb = 2 * net->depth - 2;
  FK_REQUIRE(net->d_tc_weights, "tc_pack_weights: the machine was created without the tensor-core tables");
  const TcPackDesc* d_pd = reinterpret_cast<const TcPackDesc*>((uint8_t*)net->d_tc_weights + tc_pack_offset(nb));
  tc_pack_kernel<<<nb, 256, 0, s>>>(net->d_weff, d_pd, (uint8_t*)net->d_tc_weights);
  FK_CHECK_LAUNCH();
  return 0;
}

int64_t tc_log_psi_workspace_bytes(const fk_net* net, int64_t n) {
  (void)net; (void)n;
  return 256;  // activations live in shared memory / TMEM
}

int tc_public_geometry(const fk_net* net, TcPublicGeometry* out) {
  const TcGeometry g = tc_geometry(net);
  out->P = g.P; out->p_first = g.p_first; out->npos = g.npos; out->T = g.T; out->nb = 2 * net->depth - 2;
  return g.ok ? 0 : 1;
}

int tc_log_psi(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, void* ws, int64_t ws_bytes,
               cudaStream_t s) {
  (void)ws; (void)ws_bytes;
  return tc_forward_launch(net, sigma, n, log_psi_out, nullptr, nullptr, nullptr, s);
}

int tc_forward_launch(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, uint8_t* dump,
                      uint32_t* dump_mask, float* dump_logits, cudaStream_t s, const TcWork* work, const TcxPrefix* px) {
  FK_REQUIRE(net->params_set && net->d_tc_weights, "tensor-core weights were never packed (fk_net_set_params)");
  if (n == 0) return 0;
  const TcGeometry g = tc_geometry(net);
  FK_REQUIRE(g.ok, "tensor-core engine: lattice %dx%d does not fit the shared-memory / TMEM budget", net->H, net->W);
  const int nb = 2 * net->depth - 2;
  TcArgs a;
  a.images = (const uint8_t*)net->d_tc_weights;
  a.desc = reinterpret_cast<const TcBlockDesc*>((const uint8_t*)net->d_tc_weights + tc_desc_offset(nb));
  a.sigma = sigma; a.out = log_psi_out; a.n = n;
  a.H = net->H; a.W = net->W; a.P = g.P; a.nb = nb; a.T = g.T; a.npos = g.npos; a.p_first = g.p_first; a.np = g.np;
  a.tmem_cols = g.tmem_cols; a.cst_off = (int)(g.smem_bytes - 4096); a.slots = g.slots;
  a.dump = dump; a.dump_mask = dump_mask; a.dump_logits = dump_logits;
  if (work) a.wk = *work; else a.wk = TcWork{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (px) a.px = *px; else a.px = TcxPrefix{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
  FK_REQUIRE(dump == nullptr || g.T == 1, "tensor-core gradient: lattice needs more than one M tile");
  FK_REQUIRE(px == nullptr || (g.T == 1 && !g.hreg && dump == nullptr), "prefix reuse: one-tile lattices only");
  int dev = 0, sms = 148;
  FK_CHECK_CUDA(cudaGetDevice(&dev));
  FK_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long groups = (n + g.np - 1) / g.np;
  const unsigned grid = (unsigned)std::min<long long>(groups, sms);
  const unsigned threads = 128 * g.np + 32 + 32 * TC_ISSUERS;
  auto launch = [&](auto kernel) -> int {
    FK_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    kernel<<<grid, threads, g.smem_bytes, s>>>(a);
    return 0;
  };
  int rc;
  if (dump) rc = launch(tc_forward_kernel<true, false, TC_MAX_NP>);            // (the gradient dump exists for T == 1 only)
  else if (g.hreg) rc = launch(tc_forward_kernel<false, true, 2>);
  else rc = launch(tc_forward_kernel<false, false, TC_MAX_NP>);
  if (rc) return rc;
  FK_CHECK_LAUNCH();
  return 0;
}

// ---- fp16 local energy with prefix reuse (see TcxPrefix in fk_net.cuh): the activation cache is the gradient dump ----------
// the samples' own selected log-amplitude term of every site, from the dumped logits: [b][site] float2
__global__ void tc_siteterm_kernel(const float* __restrict__ logits, const int8_t* __restrict__ sigma, long long B, int H, int W, int P,
                                   float* __restrict__ siteterm) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (e >= B * HW) return;
  const long long b = e / HW;
  const int site = (int)(e - b * HW), r = site / W, c = site - r * W;
  const float4 l4 = reinterpret_cast<const float4*>(logits)[b * 128 + r * P + c];
  const float x = 2.f * l4.x, y = 2.f * l4.y;
  const float m = fmaxf(x, y);
  const float half_lse = 0.5f * (m + logf(expf(x - m) + expf(y - m)));
  const bool up = sigma[e] > 0;
  reinterpret_cast<float2*>(siteterm)[e] = make_float2((up ? l4.x : l4.y) - half_lse, up ? l4.z : l4.w);
}

int tc_prefix_supported(const fk_net* net) {
  if (!tc_supported(net)) return 0;
  const TcGeometry g = tc_geometry(net);
  return g.ok && g.T == 1 && !g.hreg && xp_geometry_ok(net) ? 1 : 0;
}

struct TcPrefixLayout { size_t dump, mask, logits, rowcum, tiles_ws, total, stride; };
static TcPrefixLayout tc_prefix_layout(const fk_net* net, int64_t B, int64_t cap) {
  const TcGeometry g = tc_geometry(net);
  const int nb = 2 * net->depth - 2;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  TcPrefixLayout L;
  L.stride = (size_t)(nb * TC_DUMP_TENSORS + 1) * 64 * g.npos;
  size_t o = 0;
  L.dump = o; o = al(o + (size_t)B * L.stride);
  L.mask = o; o = al(o + (size_t)B * nb * TC_DUMP_TENSORS * 128 * 4);
  L.logits = o; o = al(o + (size_t)B * 128 * 16);
  L.rowcum = o; o = al(o + (size_t)B * net->sites * 8);
  L.tiles_ws = o; o = al(o + (size_t)xp_tiles_workspace_bytes(cap));
  L.total = o;
  return L;
}

int64_t tc_prefix_workspace_bytes(const fk_net* net, int64_t B, int64_t cap) { return (int64_t)tc_prefix_layout(net, B, cap).total; }

int tc_local_energy_prefix(fk_net* net, const int8_t* sigma, int64_t B, int64_t cap, const TcWork* work, void* ws, int64_t ws_bytes,
                           cudaStream_t s) {
  FK_REQUIRE(tc_prefix_supported(net), "fp16 prefix reuse: machine outside the envelope");
  const TcPrefixLayout L = tc_prefix_layout(net, B, cap);
  FK_REQUIRE((int64_t)L.total <= ws_bytes, "fp16 prefix reuse: workspace too small (%lld < %zu)", (long long)ws_bytes, L.total);
  uint8_t* base = (uint8_t*)ws;
  const TcGeometry g = tc_geometry(net);
  float* rowcum = reinterpret_cast<float*>(base + L.rowcum);
  // dump pass = the gradient's dump-mode forward over the samples (log psi of the samples, activation tiles, logits)
  if (tc_forward_launch(net, sigma, B, const_cast<float*>(work->logpsi0), base + L.dump, reinterpret_cast<uint32_t*>(base + L.mask),
                        reinterpret_cast<float*>(base + L.logits), s))
    return 1;
  tc_siteterm_kernel<<<(unsigned)((B * net->sites + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float*>(base + L.logits), sigma, B,
                                                                               net->H, net->W, g.P, rowcum);
  FK_CHECK_LAUNCH();
  const int2* tiles = nullptr;
  const long long* n_tiles = nullptr;
  if (xp_build_tiles(net, work, cap, base + L.tiles_ws, &tiles, &n_tiles, s)) return 1;
  TcxPrefix pt = {tiles, n_tiles, base + L.dump, rowcum, nullptr, nullptr, (long long)L.stride, xp_rcap(net)};
  return tc_forward_launch(net, sigma, cap, nullptr, nullptr, nullptr, nullptr, s, work, &pt);
}

}  // namespace fk

#ifdef FK_TC_TRACE
extern "C" int fk_tc_trace_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, fk::fk_tc_trace_buf, sizeof(long long) * 3 * 40 * 16);
}
#endif
