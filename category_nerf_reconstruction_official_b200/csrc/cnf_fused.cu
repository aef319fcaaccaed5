// Fused tcgen05 training step (placeholder until the kernel lands in this file).
#include "cnf_common.cuh"
using namespace cnf;
extern "C" size_t cnf_train_step_workspace_bytes(int, int, int, int, int) { return 256; }
extern "C" int cnf_train_step(int, int, int, int, int, const float*, const float*, const float*, const float*,
                              const float*, const int32_t*, const float*, const float*, const float*, const float*,
                              const uint8_t*, const uint8_t*, const int32_t*, const int32_t*, float, float, float, int,
                              float*, float*, float*, float*, float*, void*, size_t, cnf_stream_t) {
  set_error("cnf_train_step: not built yet");
  return CNF_ERR_UNSUPPORTED;
}
