// tensor-core engine placeholder (replaced by the tcgen05 fused network kernel)
#include "fk_net.cuh"
namespace fk {
int tc_supported(const fk_net* net) { (void)net; return 0; }
int tc_pack_weights(fk_net* net, cudaStream_t s) { (void)net; (void)s; return 0; }
int64_t tc_log_psi_workspace_bytes(const fk_net* net, int64_t n) { (void)net; (void)n; return 256; }
int tc_log_psi(fk_net* net, const int8_t* sigma, int64_t n, float* out, void* ws, int64_t ws_bytes, cudaStream_t s) {
  (void)net; (void)sigma; (void)n; (void)out; (void)ws; (void)ws_bytes; (void)s;
  set_error("tensor-core engine not built");
  return 1;
}
}  // namespace fk
