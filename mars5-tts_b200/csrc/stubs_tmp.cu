// TEMPORARY: placeholders so the library links while the pipeline files are being written.
#include "ctx.h"
extern "C" {
int m5_ar_generate(m5_ctx* ctx, int32_t, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const m5_ar_cfg*, int32_t, const float*, int32_t, uint64_t, const int64_t*, int32_t*, int32_t*, int32_t*, float*, int32_t) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_nar_infer(m5_ctx* ctx, int32_t, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const m5_nar_cfg*, int32_t, const int32_t*, const float*, uint64_t, const int64_t*, int32_t*) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_nar_forward(m5_ctx* ctx, int32_t, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*, int32_t, int32_t, int32_t, int32_t, float*) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_ar_forward(m5_ctx* ctx, int32_t, const int32_t*, const int32_t*, const int32_t*, const int32_t*, int32_t, float*) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_vocode(m5_ctx* ctx, int32_t, const int32_t*, const int32_t*, int32_t, int32_t, float*) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_dbg_skinny(m5_ctx* ctx, const void*, const void*, int32_t, int32_t, int32_t, float*, void*, int32_t, int32_t, int32_t) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_dbg_sample(m5_ctx* ctx, const float*, int32_t, int32_t, const m5_ar_cfg*, int32_t, const int32_t*, int32_t, const int32_t*, const int32_t*, const float*, uint64_t, int32_t*, float*) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_dbg_posterior(m5_ctx* ctx, const float*, const float*, int32_t, int32_t, int32_t, float, float, const int32_t*, const int32_t*, const uint8_t*, const float*, const float*, uint64_t, int32_t*) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
int m5_dbg_istft(m5_ctx* ctx, const float*, int32_t, const int32_t*, float*) { return ctx->fail(M5_ERR_STATE, "not built yet"); }
}
