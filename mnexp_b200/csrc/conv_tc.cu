// Title Conv1D as an implicit GEMM on tcgen05/TMEM (bf16 in, fp32 accumulate) — placeholder until the
// tensor-core kernel lands; the fp32 path (LSTUR_PREC_FP32) is complete without it.
#include "common.cuh"

struct lstur_plan;

extern "C" int lstur_conv_tc_available(void) { return 0; }

extern "C" int lstur_news_encoder_tc_fwd_internal(const lstur_plan*, const lstur_weights*, void*, int, unsigned,
                                                  cudaStream_t) {
  lstur::set_error("tensor-core news encoder not built");
  return LSTUR_ERR_UNSUPPORTED;
}
extern "C" int lstur_news_encoder_tc_bwd_internal(const lstur_plan*, const lstur_weights*, void*, float*, cudaStream_t) {
  lstur::set_error("tensor-core news encoder not built");
  return LSTUR_ERR_UNSUPPORTED;
}
