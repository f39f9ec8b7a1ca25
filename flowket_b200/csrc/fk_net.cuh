// Machine handle: the layer program of an autoregressive machine + its device-resident weights.
#pragma once
#include "fk_common.cuh"

struct fk_net {
  int kind, H, W, depth, C, k, max_dil, flags;
  int sites;
  std::vector<fk::ConvOp> ops;
  std::vector<fk::BufferInfo> bufs;   // virtual buffers (unique per producer)
  int n_phys;                         // physical slots in inference mode (liveness-based reuse)
  int phys_channels;                  // channel count of a physical slot (max over buffers)
  int in_buf, logits_buf;
  int64_t num_params;                 // raw trainable parameters
  int64_t num_eff;                    // effective weights + biases
  float* d_params;                    // raw parameters            [num_params]
  float* d_weff;                      // effective weights/biases  [num_eff]   w: [t][ci][co]
  float* d_weffT;                     // transposed kernels        [num_eff]   w: [t][co][ci]
  void* d_optable;                    // device copy of the per-op parameter mapping
  float* d_wn_dir;                    // v / |v| of the weight-normalised kernels (weff layout); 2-D machines
  float* d_wn_coef;                   // [op][2][64]: s / |v| and d w / d g over v_hat per output channel
  int64_t train_floats_per_cfg;       // sum over virtual buffers of channels * sites
  // tensor-core engine (ConvNetAutoregressive2D only): bf16 UMMA-canonical weight image
  void* d_tc_weights;
  int64_t tc_weight_bytes;
  void* d_tc_bwd;                     // transposed fp16 weight images of the tensor-core backward
  void* d_tc_exact;                   // (hi, lo) fp16 weight images of the contract-accuracy tensor-core engine
  bool params_set;
};

namespace fk {

// runs the layer program on `n` configurations.  `bufs[v]` = device pointer of virtual buffer v.
int run_forward(fk_net* net, const int8_t* sigma, int64_t n, float* const* buf_ptrs, cudaStream_t s);
// fills `ptrs` (size net->bufs.size()) for inference mode (physical slots) from a workspace base
void assign_infer_buffers(const fk_net* net, float* base, int64_t n, std::vector<float*>& ptrs);
int64_t infer_floats_per_cfg(const fk_net* net);

int launch_lncosh(const float* pre, float* out, int cout, long long npos, cudaStream_t s);

// tensor-core engine (fk_tc.cu)
int tc_supported(const fk_net* net);
int tc_pack_weights(fk_net* net, cudaStream_t s);
int64_t tc_log_psi_workspace_bytes(const fk_net* net, int64_t n);
int tc_log_psi(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, void* ws, int64_t ws_bytes,
               cudaStream_t s);

struct TcPublicGeometry { int P, p_first, npos, T, nb; };
int tc_public_geometry(const fk_net* net, TcPublicGeometry* out);

// Optional work list of the tensor-core forward kernels (the local-energy path): a work item is a connected
// configuration given as (sample, flipped sites) -- the kernel applies the flips while it loads the sample's spins, so the
// connected configurations never exist in memory -- and the last epilogue turns log psi into the local-energy term
// mel * exp(log psi' - log psi(sample)) and adds it to the sample's accumulator (operator.py:20-42).
struct TcWorkItem { int32_t sample; uint16_t site_a, site_b; };   // site 0xffff = none
struct TcWork {
  const long long* n_dev;    // number of items, read on the device (no host round trip); nullptr: the host-side n
  const TcWorkItem* items;
  const float* logpsi0;      // [B] float2: log psi of the samples
  const float* mel;          // [items] matrix elements
  double* eloc;              // [B] double2 accumulators (atomicAdd)
};
// ---- prefix reuse (dependency analysis pinned by oracle/prefix_reuse.py) ----------------------------
// A connected configuration equals its sample on every lattice row above the first flipped site (row r0), and every
// convolution of the machine looks up and sideways only, so rows >= r0 can be recomputed from the new spins plus a halo
// taken from the SAMPLE's own activations: rows r0-2, r0-1 of each block's vertical input (= relu(v') or the residual
// sum of the previous block) and of its concat tensor, row r0-1 of relu(v'); log psi(sigma') - log psi(sigma) = the sum over
// the recomputed sites of (new selected log-amplitude term - the sample's term): the rows above r0 cancel exactly.
//   dump pass   the ordinary forward over the samples also writes, per block, the (hi, lo) tiles of relu(v'), of the
//               residual sum and of the concat tensor to `dump` (the cache) and the selected term of every site to `siteterm`;
//   tile pass   a work item is a TILE holding one or two row-trimmed configurations ("segments"): segment A occupies tile
//               rows [0, kA) with its halo in the two rows above the MMA range (written by otherwise idle threads),
//               segment B tile rows [kA + 2, kA + 2 + kB) with its halo in rows kA, kA + 1 (written by the threads that own
//               those positions instead of their MMA results).  The MMA issue does not change at all -- taps are relative
//               offsets, a segment is just a translated lattice -- so two configurations share one M = 128 tile.
struct TcxPrefix {
  const int2* tiles;            // (work-list index of segment A, of segment B or -1); nullptr: not the tile pass
  const long long* n_tiles;     // device-side count
  const uint8_t* cache;         // tile pass: the samples' activation cache
  const float* rowcum;          // tile pass: [sample][sites] float2, the sample's own selected log-amplitude term of every site
                                // (log psi(sigma') - log psi(sigma) = sum over the recomputed sites of (new term - sample's term))
  uint8_t* dump;                // dump pass: cache to write (configuration i = sample i); nullptr otherwise
  float* siteterm;              // dump pass: [sample][sites] float2
  long long cache_stride;       // bytes per sample = nb * 3 tensors * 2 (hi, lo) * 64 * npos
  int rcap;                     // tile rows that lie completely inside the MMA range
};

// tiles from a device work list (fk_tc_exact.cu): row classes -> greedy pairing -> (cfgA, cfgB) list; everything stays on the device
int xp_rcap(const fk_net* net);                      // tile rows completely inside the MMA range
int xp_geometry_ok(const fk_net* net);               // one M tile, enough idle positions for the top halo
int64_t xp_tiles_workspace_bytes(int64_t cap);
int xp_build_tiles(const fk_net* net, const TcWork* work, int64_t cap, void* ws, const int2** tiles_out, const long long** n_tiles_out,
                   cudaStream_t s);

int tc_forward_launch(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, uint8_t* dump,
                      uint32_t* dump_mask, float* dump_logits, cudaStream_t s, const TcWork* work = nullptr,
                      const TcxPrefix* px = nullptr);
// fp16 tensor-core local energy with prefix reuse (fk_tc.cu)
int tc_prefix_supported(const fk_net* net);
int64_t tc_prefix_workspace_bytes(const fk_net* net, int64_t B, int64_t cap);
int tc_local_energy_prefix(fk_net* net, const int8_t* sigma, int64_t B, int64_t cap, const TcWork* work, void* ws, int64_t ws_bytes,
                           cudaStream_t s);
int tc_prepare(fk_net* net);        // allocations + wiring tables (fk_net_create)

// contract-accuracy tensor-core engine (fk_tc_exact.cu)
int tcx_supported(const fk_net* net);
int tcx_prepare(fk_net* net);
int tcx_pack_weights(fk_net* net, cudaStream_t s);
int tcx_log_psi(fk_net* net, const int8_t* sigma, int64_t n, float* log_psi_out, cudaStream_t s, const TcWork* work = nullptr);
// local energy with prefix reuse (row-trimmed connected configurations packed two per tile, fk_tc_exact.cu)
int tcx_prefix_supported(const fk_net* net);
int64_t tcx_prefix_workspace_bytes(const fk_net* net, int64_t B, int64_t cap);
int tcx_local_energy_prefix(fk_net* net, const int8_t* sigma, int64_t B, int64_t cap, const TcWork* work, void* ws, int64_t ws_bytes,
                            cudaStream_t s);

// tensor-core gradient (fk_tc_grad.cu)
int tc_grad_supported(const fk_net* net);
int tc_grad_pack_weights(fk_net* net, cudaStream_t s);
int tc_grad_prepare(fk_net* net);   // allocations + tables (fk_net_create)
int64_t tc_grad_workspace_bytes(const fk_net* net, int64_t B);
int tc_grad_weighted(fk_net* net, const int8_t* sigma, const float* y, int64_t B, float* grad_out, void* ws,
                     int64_t ws_bytes, cudaStream_t s);
int64_t grad_transform_launch(fk_net* net, const float* geff, float* graw, cudaStream_t s);
// rows: geff [m][num_eff] -> graw [m][num_params]
int grad_transform_rows_launch(fk_net* net, const float* geff, float* graw, int64_t m, cudaStream_t s);
int64_t tc_grad_per_sample_workspace_bytes(const fk_net* net, int64_t B);
int tc_grad_per_sample(fk_net* net, const int8_t* sigma, int64_t B, float* O_re, float* O_im, void* ws, int64_t ws_bytes,
                       cudaStream_t s);
// Jacobian rows straight into the panel-major bf16 operand of the sample-space Gram (see fk_jacobian_rows_tc)
int64_t tc_jacobian_rows_workspace_bytes(const fk_net* net, int64_t B);
int tc_jacobian_rows(fk_net* net, const int8_t* sigma, int64_t B, void* X, int64_t rld, int64_t row_re, int64_t row_im,
                     void* ws, int64_t ws_bytes, cudaStream_t s);

// tensor-core sampler (fk_tc_sample.cu)
int64_t tc_sample_workspace_bytes(const fk_net* net, int64_t B);
int tc_sample(fk_net* net, const double* uniforms, uint64_t seed, int64_t sample_offset, int64_t B, int8_t* sigma_out,
              float* p0_out, void* ws, int64_t ws_bytes, cudaStream_t s);

}  // namespace fk
