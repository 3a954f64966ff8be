// placeholder, replaced below
#include "mv_common.cuh"
mv_status mv_match_tc_launch(mv_ctx* ctx, const mv_match_params*, int, int, int, const int32_t*, const int32_t*,
                             const int8_t*, const int32_t*, const float*, const int32_t*, const int32_t*,
                             int32_t*, float*) {
  snprintf(ctx->err, sizeof(ctx->err), "tensor-core matcher not built");
  return MV_ERR_BAD_ARG;
}
