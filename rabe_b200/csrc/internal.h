// Private seam between the host policy layer (host_capi.cpp, no CUDA) and the engine (engine.cu).
#pragma once
#include <stdint.h>
#include "../../include/rabe_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
// terms: n_terms x {coef, x, e, pad} uint32; leaf_offs: n_leaves + 1 offsets into terms
int rb_share_plan_create_raw(rb_ctx*, const uint32_t* terms, uint32_t n_terms, const uint32_t* leaf_offs, uint32_t n_leaves,
                             uint32_t n_coefs, rb_share_plan** out);
int rb_lagrange_raw(rb_ctx*, const uint32_t* terms, uint32_t n_terms, const uint32_t* leaf_offs, uint32_t n_leaves, uint8_t* out);
#ifdef __cplusplus
}
#endif
