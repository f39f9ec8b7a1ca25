// Shared declarations of the flowket_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/flowket_b200.h"

namespace fk {

void set_error(const char* fmt, ...);

#define FK_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      fk::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

#define FK_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      fk::set_error(__VA_ARGS__);  \
      return 1;                    \
    }                              \
  } while (0)

void count_launch();
#define FK_CHECK_LAUNCH()               \
  do {                                  \
    fk::count_launch();                 \
    FK_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

constexpr int ACT_NONE = 0;
constexpr int ACT_RELU = 1;
constexpr int ACT_LNCOSH = 2;  // complex pairs (channel c, channel c + cout/2)

constexpr int MAX_TAPS = 9;

// One convolution of the layer program.  Activations are NHWC fp32 with a per-buffer channel count;
// a tap t reads the input at spatial offset (dh[t], dw[t]) (zero outside the lattice), which folds
// ZeroPadding2D, DownShift/RightShift and dilation into the gather.
struct ConvOp {
  int in_buf, cin;
  int out_buf, out_coff, cout;
  int out2_buf;   // -1, or receives relu(z) (z = conv + bias, before the residual add)
  int res_buf;    // -1, or added to z before the activation
  int pre_buf;    // -1, or receives z (needed by the lncosh backward)
  int act;
  int ntaps;
  int dh[MAX_TAPS], dw[MAX_TAPS];
  int64_t w_off, b_off;  // offsets into the effective-weight buffer: [ntaps*cin][cout], [cout]
  // mapping to the raw (trainable) parameter vector
  int64_t p_kernel, p_bias, p_g;  // p_g = -1 without weight normalisation
  int64_t p_kernel_imag, p_bias_imag;  // complex nets only (-1 otherwise)
  int raw_cin, raw_cout;           // shape of the raw kernel (complex nets: half of cin/cout)
};

struct BufferInfo {
  int channels;
  int phys;  // physical slot in inference mode
};

struct ConvLaunch {
  const float* in; int in_cs; int cin;
  float* out; int out_cs; int out_coff; int cout;
  float* out2; int out2_cs;
  const float* res; int res_cs;
  float* pre; int pre_cs;
  const float* w; const float* bias;
  int ntaps; int dh[MAX_TAPS]; int dw[MAX_TAPS];
  int H, W; long long npos;
  int act; int accumulate;  // accumulate: out += result (backward-data); no bias/act then
};

int launch_conv(const ConvLaunch& a, cudaStream_t s);

// dz = g_out * act'(out)  (+ g_out2 * [out2 > 0]);  g_res += g_out * act'(out)
int launch_dz(const float* g_out, int g_cs, int g_coff, const float* out, int out_cs, int out_coff,
              const float* g_out2, const float* out2, float* g_res, const float* pre, int cout, int act,
              float* dz, long long npos, cudaStream_t s);

// dW[t][ci][co] (+)= sum_p in(p + delta_t)[ci] * dz(p)[co];  db[co] (+)= sum_p dz(p)[co]
// per_sample = 0: one accumulated result (atomics);  per_sample = 1: one result per sample, written to
// out + b * out_stride (no atomics).
int launch_dw(const float* in, int in_cs, int cin, const float* dz, int cout, int ntaps, const int* dh,
              const int* dw, int H, int W, long long n, float* dW, float* db, int per_sample,
              long long out_stride, cudaStream_t s);

int launch_sigma_to_float(const int8_t* sigma, float* out, int channels, long long n_elems, cudaStream_t s);
int launch_head(const float* logits, const int8_t* sigma, int sites, long long n, float* log_psi,
                float* cond_log_probs, cudaStream_t s);
int launch_head_backward(const float* logits, const int8_t* sigma, int sites, long long n, const float* coef_re,
                         const float* coef_im, float* g_logits, cudaStream_t s);

}  // namespace fk
