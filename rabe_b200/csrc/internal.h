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
// test hooks of the six-lane pairing layer (wide.cuh): the wide accumulator and single Fq12 operations
int rb_dbg_wide_dot(rb_ctx*, const uint8_t* xs, const uint8_t* ys, int K, size_t n, uint8_t* out);
int rb_dbg_w6_op(rb_ctx*, int op, int arg, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);
// test hook of the dedicated Montgomery squaring (fp.cuh fe_sqr): mode 0 Fq, 1 Fr, 2 Fq with an unreduced operand a + b
int rb_dbg_fq_sqr(rb_ctx*, const uint8_t* a, const uint8_t* b, int mode, size_t n, uint8_t* out);
#ifdef __cplusplus
}
#endif
