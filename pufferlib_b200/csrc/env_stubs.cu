// env_stubs.cu -- placeholders for env kinds whose kernels live in their own translation units.
#include "env_common.cuh"
#ifndef PB_HAVE_BREAKOUT
int pb_breakout_create(pb_env*) { pb_set_error("breakout: not built into this library"); return PB_ERR_UNSUPPORTED; }
#endif
#ifndef PB_HAVE_SNAKE
int pb_snake_create(pb_env*) { pb_set_error("snake: not built into this library"); return PB_ERR_UNSUPPORTED; }
#endif
#ifndef PB_HAVE_PONG
int pb_pong_create(pb_env*) { pb_set_error("pong: not built into this library"); return PB_ERR_UNSUPPORTED; }
#endif
#ifndef PB_HAVE_IMAGE
int pb_minibatch_gather_tma(const void*, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                            cudaStream_t) { pb_set_error("TMA gather: not built"); return PB_ERR_UNSUPPORTED; }
extern "C" int pb_image_pack(const void*, int64_t, const void*, int64_t, void*, int64_t, const uint8_t*, int64_t, int64_t,
                             int32_t, void*) { pb_set_error("pb_image_pack: not built"); return PB_ERR_UNSUPPORTED; }
#endif
#ifndef PB_HAVE_SAMPLE
extern "C" int pb_sample_logits(const float*, int64_t, int32_t, uint64_t, uint64_t, int64_t*, float*, float*, const float*,
                                float*, float*, int64_t*, void*) { pb_set_error("pb_sample_logits: not built"); return PB_ERR_UNSUPPORTED; }
#endif
