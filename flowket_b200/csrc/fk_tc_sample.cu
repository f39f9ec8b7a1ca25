// Tensor-core incremental sampler for ConvNetAutoregressive2D (C = 32, k = 3): the same cached exact ancestral
// sampling as fk_sample.cu (every (layer, site) activation computed once), with the per-site convolutions issued as
// tcgen05 MMAs: M = 128 samples of the CTA, N = 32/16 output channels, K = 16 per instruction.
//
//   * activation caches are fp16 tiles [channel group of 8][128 samples][8 ch] (8 KB, already in the UMMA canonical
//     K-major layout) in CTA-private global memory; the taps a step needs are fetched with cp.async.bulk into shared
//     memory, tiles produced by the step stay in shared memory for the next conv and are written back to the cache;
//   * weights: the per-block fp16 images of the fused forward kernel (fk_tc.cu), double-buffered;
//   * one thread issues the MMAs, 128 threads (one TMEM lane = one sample each) run the epilogues; the last block's
//     epilogue fuses head -> log-space normalisation -> explicit-uniform draw (deepar/samplers/autoregressive.py:37-44);
//   * schedule per row: for every column [last block + head + draw, blocks 0..nb-2 horizontal], then the vertical stack
//     of the row for all blocks.
// Numerics: fp16 operands / fp32 accumulation; spins agree with the fp32 sampler except where |p0 - u| is within the
// fp16 error of p0 (statistically exact sampling from the fp16-evaluated network; tolerance in tests/test_gpu_tc.py).
#include <algorithm>

#include "fk_net.cuh"
#include "fk_tc_common.cuh"

namespace fk {

// Philox4x32-10 (same stream as fk_sample.cu)
__device__ __forceinline__ double tcs_philox_uniform(uint64_t seed, uint64_t sample, uint32_t site) {
  uint32_t c0 = (uint32_t)sample, c1 = (uint32_t)(sample >> 32), c2 = site, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint64_t hi = c0 >> 5, lo = c1 >> 6;
  return (double)((hi << 26) | lo) * (1.0 / 9007199254740992.0);
}

#ifdef FK_TS_TRACE
// clock64 timeline of CTA 0 at site (5, 5): [block step kb][slot]  (tools/sampler_trace.py)
__device__ long long fk_ts_trace_buf[40 * 16];
#define TSTRACE(slot) do { if (blockIdx.x == 0 && tid == 0 && i == 5 && j == 5 && kb < 40) fk_ts_trace_buf[kb * 16 + (slot)] = clock64(); } while (0)
#else
#define TSTRACE(slot) do {} while (0)
#endif

constexpr int TS_TILE = 8192;       // bytes of one activation tile
constexpr int TS_NLOAD = 12;        // loaded-tile slots

struct TcSampleArgs {
  const uint8_t* images;   // nb weight images (fk_tc.cu layout)
  uint8_t* cache;          // CTA-private tile caches
  long long cache_tiles_per_cta;
  int H, W, nb;
  const double* uniforms;
  uint64_t seed;
  long long sample_offset, B;
  int8_t* sigma_out;
  float* p0_out;
};

// cache tile index (per CTA): block b owns 8W tiles: vin[3][W] | hin[W] | a[W] | c[3][W]
__device__ __forceinline__ long long ts_vin(int W, int b, int slot, int col) { return (long long)b * 8 * W + slot * W + col; }
__device__ __forceinline__ long long ts_hin(int W, int b, int col) { return (long long)b * 8 * W + 3 * W + col; }
__device__ __forceinline__ long long ts_a(int W, int b, int col) { return (long long)b * 8 * W + 4 * W + col; }
__device__ __forceinline__ long long ts_c(int W, int b, int slot, int col) { return (long long)b * 8 * W + 5 * W + slot * W + col; }

__global__ void __launch_bounds__(128, 1) tc_sample_kernel(TcSampleArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  // MMA issue: whole warp 0 walks the (CTA-uniform) control flow and one elected lane issues -- under `if (tid == 0)`
  // the compiler moves every descriptor through a per-thread R2UR loop (~100 cycles per tcgen05.mma)
  const bool mma_warp = __shfl_sync(0xffffffffu, warp, 0) == 0;
  uint8_t* wbuf = smem;                                   // 2 weight images
  uint8_t* ltile = smem + 2 * IMG_CORE_BYTES;                  // TS_NLOAD loaded tiles
  uint8_t* xc = ltile + TS_NLOAD * TS_TILE;               // x1, then the concat tensor (in place)
  uint8_t* hring = xc + TS_TILE;                          // 3 tiles: horizontal-stack tile of the current site per block parity
  uint8_t* tail = hring + 3 * TS_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);     // wfull[0..1], tfull (x/a tiles)[2], mma[3], cfull (c / vertical tiles)[4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 64);

  const uint32_t wfull0 = smem_u32(&bars[0]), tfull = smem_u32(&bars[2]), mbar = smem_u32(&bars[3]), cfull = smem_u32(&bars[4]);
  if (tid == 32) {
    mbar_init(wfull0, 1); mbar_init(wfull0 + 8, 1); mbar_init(tfull, 1); mbar_init(mbar, 1); mbar_init(cfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;

  const int H = a.H, W = a.W, nb = a.nb, sites = H * W;
  uint8_t* cache = a.cache + (size_t)blockIdx.x * a.cache_tiles_per_cta * TS_TILE;
  const long long gb = (long long)blockIdx.x * 128 + tid;   // global sample row of this thread
  const bool live = gb < a.B;

  const uint32_t idesc32 = make_idesc(32), idesc16 = make_idesc(16);
  const uint64_t adesc0 = make_desc(0, 128, 8);              // tile: channel groups 128*16 B apart, 8-row groups 128 B
  const uint64_t bdesc32 = make_desc(0, 32, 8), bdesc16 = make_desc(0, 16, 8);
  const uint32_t ltile16 = smem_u32(ltile) >> 4, xc16 = smem_u32(xc) >> 4, hring16 = smem_u32(hring) >> 4;
  constexpr uint32_t TILE16 = TS_TILE / 16, KSTEP16 = 256;

  uint32_t mma_phase = 0, tile_phase = 0, c_phase = 0;
  long long step = 0;   // weight-ring step

  // ---- helpers ---------------------------------------------------------------------------------------------------
  auto store_tile_row = [&](uint8_t* tile_smem, uint8_t* tile_gmem, const float* v) {
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      uint4 q;
      q.x = pack_h2(v[8 * cg + 0], v[8 * cg + 1]);
      q.y = pack_h2(v[8 * cg + 2], v[8 * cg + 3]);
      q.z = pack_h2(v[8 * cg + 4], v[8 * cg + 5]);
      q.w = pack_h2(v[8 * cg + 6], v[8 * cg + 7]);
      const size_t off = (size_t)(cg * 128 + tid) * 16;
      if (tile_smem) *reinterpret_cast<uint4*>(tile_smem + off) = q;
      if (tile_gmem) *reinterpret_cast<uint4*>(tile_gmem + off) = q;
    }
  };
  auto load_tile_row = [&](const uint8_t* tile_smem, float* v) {
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      const uint4 q = *reinterpret_cast<const uint4*>(tile_smem + (size_t)(cg * 128 + tid) * 16);
      unpack_h8(q, v + 8 * cg);
    }
  };
  // 2 k-steps of one tap: A = tile (16-byte units), B = weight tile
  auto mma_tap = [&](uint32_t d_col, uint32_t tile16, uint64_t bd, uint32_t wstep, uint32_t idesc, uint32_t& acc) {
    const uint64_t ad = adesc0 + (uint64_t)tile16;
    if (elect_one()) {   // (called by the whole MMA warp: the descriptors stay in uniform registers)
      umma_f16(tmem + d_col, ad, bd, idesc, acc);
      umma_f16(tmem + d_col, ad + KSTEP16, bd + wstep, idesc, 1u);
    }
    acc = 1;
  };
  auto commit_and_wait = [&]() {
    if (mma_warp && elect_one()) umma_commit(mbar);   // same elected lane as the MMAs
    mbar_wait(mbar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };
  // Epilogue writes to shared memory feed the next MMA (generic -> async proxy): fenced every phase.  Writes to the
  // global caches are only read back by cp.async.bulk one site later (horizontal stack) or one block later (vertical
  // stack), so the much more expensive global proxy fence is issued once per site / once per vertical block.
  auto end_phase = [&](bool global_fence = false) {
    fence_proxy_async();
    if (global_fence) asm volatile("fence.proxy.async.global;" ::: "memory");
    tc_fence_before();
    __syncthreads();
  };
  // weight ring: image `img` for ring step s
  auto load_weights = [&](long long s, int img) {
    const uint32_t sel = (uint32_t)(s & 1);
    mbar_expect_tx(wfull0 + 8 * sel, IMG_CORE_BYTES);
    bulk_g2s(smem_u32(wbuf + (size_t)sel * IMG_CORE_BYTES), a.images + (size_t)img * IMG_BYTES, IMG_CORE_BYTES, wfull0 + 8 * sel);
  };
  // image sequence of a row: for every column [nb-1, 0, 1, ..., nb-2], then the vertical pass [0..nb-1]
  auto image_of = [&](long long s) -> int {
    const long long per_row = (long long)W * nb + nb;
    const long long r = s % per_row;
    if (r < (long long)W * nb) {
      const int k = (int)(r % nb);
      return k == 0 ? nb - 1 : k - 1;
    }
    return (int)(r - (long long)W * nb);
  };
  const long long total_steps = (long long)H * ((long long)W * nb + nb);

  if (tid == 0) { load_weights(0, image_of(0)); }

  // ================================================================================================================
  // Loaded-tile slots of a horizontal step: 0..2 = 1x3 taps of the horizontal stack, 3 = relu(v')(i-1, j),
  // 4..11 = the eight concat-tensor taps that come from the cache (the ninth is produced by the step itself).
  // The x/a tiles of step k+1 are prefetched while step k runs its 3x3 conv; the concat tiles of step k are
  // requested at the start of the step and only awaited before the 3x3 conv.
  auto x_valid = [&](int b, int j, int t) -> bool {            // tap t of the 1x3 conv of block b at column j
    const bool last = (b == nb - 1);
    const int jc = last ? j - 1 : j;
    return jc >= 0 && (jc - 2 + t) >= 0;
  };
  auto issue_xa = [&](int i, int j, int b) -> int {            // thread 0 only; returns the number of tiles requested
    const bool last = (b == nb - 1);
    const int jc = last ? j - 1 : j;
    int n = 0;
    for (int t = 0; t < 2; ++t) n += x_valid(b, j, t) ? 1 : 0;  // tap 2 (column jc) always comes from the h ring
    if (i > 0) ++n;
    if (n == 0) return 0;
    mbar_expect_tx(tfull, (uint32_t)n * TS_TILE);
    for (int t = 0; t < 2; ++t)
      if (x_valid(b, j, t))
        bulk_g2s(smem_u32(ltile + (size_t)t * TS_TILE), cache + (size_t)ts_hin(W, b, jc - 2 + t) * TS_TILE, TS_TILE, tfull);
    if (i > 0) bulk_g2s(smem_u32(ltile + (size_t)3 * TS_TILE), cache + (size_t)ts_a(W, b, j) * TS_TILE, TS_TILE, tfull);
    return n;
  };
  auto count_xa = [&](int i, int j, int b) -> int {
    int n = 0;
    for (int t = 0; t < 2; ++t) n += x_valid(b, j, t) ? 1 : 0;
    return n + (i > 0 ? 1 : 0);
  };

  for (int i = 0; i < H; ++i) {
    bool xa_prefetched = false;
    for (int j = 0; j < W; ++j) {
      for (int kb = 0; kb < nb; ++kb, ++step) {
        const int b = kb == 0 ? nb - 1 : kb - 1;     // last block first (it produces sigma(i,j)), then blocks 0..nb-2
        const bool last = (b == nb - 1);
        const int jc = last ? j - 1 : j;             // RightShift: the last block evaluates its 1x3 conv one column to the left
        const uint32_t wsel = (uint32_t)(step & 1);
        // ---- request this step's concat tiles, (first step of a row: also its x/a tiles), prefetch the next weight image
        int nc = 0;
        int c_slot[9];
        for (int di = 0; di < 3; ++di)
          for (int dj = 0; dj < 3; ++dj) {
            const int row = i - 2 + di, col = j - 2 + dj;
            c_slot[di * 3 + dj] = -1;
            if (row < 0 || col < 0) continue;
            if (di == 2 && dj == 2) { c_slot[8] = 100; continue; }   // produced by this step (xc tile)
            c_slot[di * 3 + dj] = 4 + di * 3 + dj;
            ++nc;
          }
        const int nxa = count_xa(i, j, b);
        TSTRACE(0);
        if (tid == 0) {
          if (!xa_prefetched) issue_xa(i, j, b);
          if (nc > 0) {
            mbar_expect_tx(cfull, (uint32_t)nc * TS_TILE);
            for (int t = 0; t < 8; ++t)
              if (c_slot[t] >= 0)
                bulk_g2s(smem_u32(ltile + (size_t)c_slot[t] * TS_TILE),
                         cache + (size_t)ts_c(W, b, (i - 2 + t / 3) % 3, j - 2 + t % 3) * TS_TILE, TS_TILE, cfull);
          }
          if (step + 1 < total_steps) load_weights(step + 1, image_of(step + 1));
        }
        mbar_wait(wfull0 + 8 * wsel, (uint32_t)((step >> 1) & 1));
        if (nxa > 0) { mbar_wait(tfull, tile_phase); tile_phase ^= 1; }
        TSTRACE(1);
        const uint8_t* wimg = wbuf + (size_t)wsel * IMG_CORE_BYTES;
        const uint32_t wimg16 = smem_u32(wimg) >> 4;
        const float* bias = reinterpret_cast<const float*>(wimg + IMG_BIAS);
        uint8_t* h_out = hring + (size_t)((b + 1) % 3) * TS_TILE;
        const uint8_t* h_res = hring + (size_t)((b + 2) % 3) * TS_TILE;   // (b-1) mod 3: the pair input at this site
        const bool res2 = (b >= 2 && (b % 2) == 0 && !last);

        // ================= phase 1: 1x3 conv on the horizontal stack -> x1
        const bool have_x = jc >= 0;
        if (mma_warp && have_x) {
          tc_fence_after();
          uint32_t acc = 0;
          for (int t = 0; t < 3; ++t) {
            if (t < 2 && !x_valid(b, j, t)) continue;
            // tap 2 = column jc: the tile this CTA produced last (block b-1 at this site, or block nb-2 at the previous site)
            const uint32_t tile16 = t == 2 ? hring16 + (uint32_t)(b % 3) * TILE16 : ltile16 + (uint32_t)t * TILE16;
            mma_tap(0, tile16, bdesc32 + wimg16 + IMG_X / 16 + (uint64_t)(t * 2) * 64, 64, idesc32, acc);
          }
        }
        if (have_x) commit_and_wait();
        TSTRACE(2);
        {
          float v[32];
          if (have_x) {
            tmem_ld32(tmem + lane_sel + 0, v);
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q] + bias[32 + q], 0.f);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0.f;   // RightShift pads zeros after the activation
          }
          store_tile_row(xc, nullptr, v);
        }
        end_phase();
        TSTRACE(3);

        // ================= phase 2: 1x1 convs: x1 -> concat[0:16], DownShift(relu(v')) = a(i-1, j) -> concat[16:32]
        if (mma_warp) {
          tc_fence_after();
          uint32_t acc = 0;
          mma_tap(32, xc16, bdesc16 + wimg16 + IMG_XX / 16, 32, idesc16, acc);
          if (i > 0) {
            acc = 0;
            mma_tap(48, ltile16 + 3u * TILE16, bdesc16 + wimg16 + IMG_Y / 16, 32, idesc16, acc);
          }
        }
        commit_and_wait();
        TSTRACE(4);
        // the x/a slots are free now: prefetch the next step's x/a tiles (same row only; the vertical pass reuses the slots)
        {
          int ni = i, nj = j, nkb = kb + 1;
          if (nkb == nb) { nkb = 0; ++nj; }
          xa_prefetched = nj < W;
          if (xa_prefetched && tid == 0) issue_xa(ni, nj, nkb == 0 ? nb - 1 : nkb - 1);
        }
        {
          float v[32];
          tmem_ld32(tmem + lane_sel + 32, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q] + bias[64 + q], 0.f);
#pragma unroll
          for (int q = 16; q < 32; ++q) v[q] = fmaxf((i > 0 ? v[q] : 0.f) + bias[64 + q], 0.f);
          store_tile_row(xc, cache + (size_t)ts_c(W, b, i % 3, j) * TS_TILE, v);
        }
        end_phase();
        TSTRACE(5);

        // ================= phase 3: 3x3 conv on the concat tensor -> h'
        if (nc > 0) { mbar_wait(cfull, c_phase); c_phase ^= 1; }
        TSTRACE(6);
        if (mma_warp) {
          tc_fence_after();
          uint32_t acc = 0;
          for (int t = 0; t < 9; ++t) {
            if (c_slot[t] < 0) continue;
            const uint32_t tile16 = c_slot[t] == 100 ? xc16 : ltile16 + (uint32_t)c_slot[t] * TILE16;
            mma_tap(64, tile16, bdesc32 + wimg16 + IMG_H / 16 + (uint64_t)(t * 2) * 64, 64, idesc32, acc);
          }
        }
        commit_and_wait();
        TSTRACE(7);
        {
          float v[32];
          tmem_ld32(tmem + lane_sel + 64, v);
          if (res2) {
            float r[32];
            load_tile_row(h_res, r);
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += r[q];
          }
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q] + bias[96 + q], 0.f);
          store_tile_row(h_out, last ? nullptr : cache + (size_t)ts_hin(W, b + 1, j) * TS_TILE, v);
        }
        end_phase(kb == nb - 1);   // end of the site: publish this site's cache tiles to the async proxy
        TSTRACE(8);

        // ================= phase 4 (last block): head + normalisation + draw sigma(i,j)
        if (last) {
          const float* hb = reinterpret_cast<const float*>(wimg + IMG_HEAD_BIAS);
          if (mma_warp) {
            tc_fence_after();
            uint32_t acc = 0;
            mma_tap(96, hring16 + (uint32_t)((b + 1) % 3) * TILE16, bdesc16 + wimg16 + IMG_HEAD / 16, 32, idesc16, acc);
          }
          commit_and_wait();
          float lg[16];
          tmem_ld16(tmem + lane_sel + 96, lg);
          const float re0 = lg[0] + hb[0], re1 = lg[1] + hb[1];
          const float x = 2.f * re0, y = 2.f * re1;
          const float m = fmaxf(x, y);
          const float lse = m + logf(expf(x - m) + expf(y - m));
          const float p0 = expf(2.f * (re0 - 0.5f * lse));
          const int site = i * W + j;
          double u = 2.0;
          if (live) u = a.uniforms ? a.uniforms[gb * sites + site] : tcs_philox_uniform(a.seed, (uint64_t)(a.sample_offset + gb), (uint32_t)site);
          const float sg = ((double)p0 > u) ? 1.f : -1.f;
          if (live) {
            a.sigma_out[gb * sites + site] = (int8_t)sg;
            if (a.p0_out) a.p0_out[gb * sites + site] = p0;
          }
          // sigma as a 32-channel tile (channel 0): input of block 0 in both stacks
          float v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0.f;
          v[0] = sg;
          store_tile_row(hring, cache + (size_t)ts_hin(W, 0, j) * TS_TILE, v);       // ring slot 0 = hin of block 0
          store_tile_row(nullptr, cache + (size_t)ts_vin(W, 0, i % 3, j) * TS_TILE, v);
          end_phase();
        }
      }
    }
    // ================= vertical stack of row i: all blocks, all columns.  Rolling window: slot = 4*row_tap + (column & 3);
    // column j needs columns j-1..j+1, column j+2 and the residual tile of column j+1 are prefetched meanwhile.
    for (int b = 0; b < nb; ++b, ++step) {
      const uint32_t wsel = (uint32_t)(step & 1);
      const bool res2 = (b >= 2 && (b % 2) == 0 && b != nb - 1);
      auto issue_col = [&](int col, int res_col) {   // thread 0: the three row taps of `col` (+ residual tile of res_col)
        int n = 0;
        for (int di = 0; di < 3; ++di) n += (i - 2 + di >= 0 && col >= 0 && col < W) ? 1 : 0;
        if (res2 && res_col >= 0 && res_col < W) ++n;
        if (n == 0) return;
        mbar_expect_tx(cfull, (uint32_t)n * TS_TILE);
        if (col >= 0 && col < W)
          for (int di = 0; di < 3; ++di)
            if (i - 2 + di >= 0)
              bulk_g2s(smem_u32(ltile + (size_t)(di * 4 + (col & 3)) * TS_TILE),
                       cache + (size_t)ts_vin(W, b, (i - 2 + di) % 3, col) * TS_TILE, TS_TILE, cfull);
        if (res2 && res_col >= 0 && res_col < W)
          bulk_g2s(smem_u32(hring + (size_t)(1 + (res_col & 1)) * TS_TILE), cache + (size_t)ts_vin(W, b - 1, i % 3, res_col) * TS_TILE,
                   TS_TILE, cfull);
      };
      if (tid == 0) {
        if (step + 1 < total_steps) load_weights(step + 1, image_of(step + 1));
        issue_col(0, 0);        // group "column 0": column 0 tiles + residual of column 0
      }
      mbar_wait(wfull0 + 8 * wsel, (uint32_t)((step >> 1) & 1));
      mbar_wait(cfull, c_phase); c_phase ^= 1;
      if (tid == 0 && W > 1) issue_col(1, -1);   // column 1 is needed by column 0's right tap
      if (W > 1) { mbar_wait(cfull, c_phase); c_phase ^= 1; }
      const uint8_t* wimg = wbuf + (size_t)wsel * IMG_CORE_BYTES;
      const uint32_t wimg16 = smem_u32(wimg) >> 4;
      const float* bias = reinterpret_cast<const float*>(wimg + IMG_BIAS);
      for (int j = 0; j < W; ++j) {
        // prefetch: column j+2 (right tap of column j+1) and the residual tile of column j+1
        const bool pre = (j + 2 < W) || (res2 && j + 1 < W);
        if (tid == 0 && pre) issue_col(j + 2 < W ? j + 2 : -1, j + 1);
        if (mma_warp) {
          tc_fence_after();
          uint32_t acc = 0;
          for (int di = 0; di < 3; ++di)
            for (int dj = 0; dj < 3; ++dj) {
              const int row = i - 2 + di, col = j - 1 + dj;
              if (row < 0 || col < 0 || col >= W) continue;
              mma_tap(0, ltile16 + (uint32_t)(di * 4 + (col & 3)) * TILE16,
                      bdesc32 + wimg16 + IMG_V / 16 + (uint64_t)((di * 3 + dj) * 2) * 64, 64, idesc32, acc);
            }
        }
        commit_and_wait();
        {
          float v[32];
          tmem_ld32(tmem + lane_sel + 0, v);
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] += bias[q];
          if (b + 1 < nb) {
            float r[32];
            if (res2) {
              load_tile_row(hring + (size_t)(1 + (j & 1)) * TS_TILE, r);
#pragma unroll
              for (int q = 0; q < 32; ++q) r[q] = fmaxf(r[q] + v[q], 0.f);
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q) r[q] = fmaxf(v[q], 0.f);
            }
            store_tile_row(nullptr, cache + (size_t)ts_vin(W, b + 1, i % 3, j) * TS_TILE, r);
          }
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
          store_tile_row(nullptr, cache + (size_t)ts_a(W, b, j) * TS_TILE, v);
        }
        end_phase(j == W - 1);     // the next block reads this block's output tiles
        if (pre) { mbar_wait(cfull, c_phase); c_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

constexpr size_t TS_SMEM = 2 * (size_t)IMG_CORE_BYTES + (size_t)(TS_NLOAD + 1 + 3) * TS_TILE + 128;

int64_t tc_sample_workspace_bytes(const fk_net* net, int64_t B) {
  const int nb = 2 * net->depth - 2;
  const int64_t ctas = (B + 127) / 128;
  return ctas * (int64_t)nb * 8 * net->W * TS_TILE + 256;
}

int tc_sample(fk_net* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B, int8_t* sigma_out,
              float* p0_out, void* ws, int64_t ws_bytes, cudaStream_t s) {
  FK_REQUIRE(net->params_set && net->d_tc_weights, "tensor-core weights were never packed (fk_net_set_params)");
  FK_REQUIRE(ws_bytes >= tc_sample_workspace_bytes(net, B), "fk_sample (tensor-core engine): workspace too small");
  if (B == 0) return 0;
  TcSampleArgs a;
  a.images = (const uint8_t*)net->d_tc_weights;
  a.cache = (uint8_t*)ws;
  a.H = net->H; a.W = net->W; a.nb = 2 * net->depth - 2;
  a.cache_tiles_per_cta = (long long)a.nb * 8 * net->W;
  a.uniforms = uniforms; a.seed = seed; a.sample_offset = sample_offset; a.B = B;
  a.sigma_out = sigma_out; a.p0_out = p0_out;
  FK_CHECK_CUDA(cudaFuncSetAttribute(tc_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM));
  tc_sample_kernel<<<(unsigned)((B + 127) / 128), 128, TS_SMEM, s>>>(a);
  FK_CHECK_LAUNCH();
  return 0;
}

}  // namespace fk

#ifdef FK_TS_TRACE
extern "C" int fk_ts_trace_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, fk::fk_ts_trace_buf, sizeof(long long) * 40 * 16);
}
#endif
