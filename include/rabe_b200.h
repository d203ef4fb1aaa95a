/* rabe_b200 -- C ABI of the B200-native batched ABE engine (librabe_b200.so).
 *
 * This is the drop-in boundary for the hot path of Fraunhofer-AISEC/rabe: the per-attribute /
 * per-MSP-row G1/G2 scalar multiplications, Gt exponentiations and pairing products that rabe's
 * schemes obtain from the external crate `rabe_bn` (rabe Cargo.toml:33).  rabe has no FFI on this
 * path; the seam is the operator API of rabe_bn as used inside src/schemes/{ac17,bsw,lsw,aw11}/mod.rs
 * plus the share computation of src/utils/secretsharing/mod.rs.  Each entry point below names the
 * reference statements (file:line under /root/reference/src) it replaces.  INTEGRATION.md shows
 * the Rust `extern "C"` block a rabe maintainer would add.
 *
 * Conventions
 *  - Return value: RB_OK (0) or a negative rb_status; nothing aborts or throws.
 *  - The caller owns every data buffer.  The library only allocates opaque handles (rb_ctx,
 *    rb_table, rb_ac17_pk, rb_msp ...) that have a matching *_destroy / *_free.
 *  - Every data pointer may be a HOST pointer or a DEVICE pointer (of the context's GPU); the
 *    library detects which.  Host buffers are staged through the context's stream (the call
 *    returns after the results are back in the host buffers).  With device buffers the call only
 *    enqueues work on the context's stream; use rb_ctx_sync()/rb_ctx_status() to wait.
 *  - Element encodings (canonical, fixed): integers are 32-byte big-endian, fully reduced, NOT in
 *    Montgomery form.
 *        Fr   32 B
 *        G1   64 B   x | y            (affine; all-zero = point at infinity)
 *        G2  128 B   x.re | x.im | y.re | y.im   (affine, on the twist y^2 = x^3 + 3/(9+i))
 *        Gt  384 B   12 Fq coefficients in tower order c0.c0.re, c0.c0.im, c0.c1.re, ... c1.c2.im
 *                    for Fq12 = Fq6[w]/(w^2-v), Fq6 = Fq2[v]/(v^3-(9+i)), Fq2 = Fq[i]/(i^2+1)
 *  - All randomness is an explicit input (rabe draws it from rand::thread_rng()).
 *  - One rb_ctx per (host thread, GPU).  Calls on distinct contexts are independent.
 *  - There is no CPU fallback: without a CUDA device every entry point returns RB_ECUDA.
 *  - Input validation: every field element must be canonical (< modulus), every G1/G2 point on its
 *    curve, and every G2 point supplied by the caller in the order-r subgroup -- the checks rabe_bn
 *    performs when such values are deserialised (FieldError::NotMember -> RabeError, error.rs:60-69).
 *    A violation returns RB_ENOTMEMBER.  The G2 subgroup test (one 63-bit scalar multiplication per
 *    point) can be waived per context for inputs the caller already validated or produced with this
 *    library: rb_ctx_set_g2_subgroup_check().
 *  - Index lists (idx / offs) in host memory are range-checked (RB_EINVAL).  Lists in DEVICE memory
 *    are the caller's contract: offs non-decreasing with offs[last] <= n_idx, every idx in range.
 */
#ifndef RABE_B200_H
#define RABE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rb_status {
  RB_OK = 0,
  RB_EINVAL = -1,     /* bad argument (null pointer, zero size, unsupported window ...)            */
  RB_ENOTMEMBER = -2, /* an input is not a canonical field element / not on the curve
                         (rabe: FieldError::NotMember -> RabeError, error.rs:60-69)                */
  RB_EPOLICY = -3,    /* policy / MSP description is inconsistent (rabe panics: msp.rs:121,133)    */
  RB_ECUDA = -4,      /* CUDA runtime failure or no device                                          */
  RB_ENOMEM = -5
} rb_status;

#define RB_FR_BYTES 32
#define RB_G1_BYTES 64
#define RB_G2_BYTES 128
#define RB_GT_BYTES 384

typedef struct rb_ctx rb_ctx;
typedef struct rb_table rb_table;

const char* rb_strerror(int status);
const char* rb_version(void);

/* ---- context ------------------------------------------------------------------------------ */
int rb_ctx_create(int device, rb_ctx** out);
void rb_ctx_destroy(rb_ctx* ctx);
/* Launch on an existing CUDA stream (cudaStream_t passed as void*; NULL is the legacy default
 * stream).  rb_ctx_reset_stream() returns to the context's own non-blocking stream. */
int rb_ctx_set_stream(rb_ctx* ctx, void* cuda_stream);
int rb_ctx_reset_stream(rb_ctx* ctx);
/* The cudaStream_t (as void*) the context currently launches on -- for callers that must order their own
 * work with it (cudaStreamWaitEvent): with DEVICE buffers a call only enqueues, so inputs written on another
 * stream must be complete, and outputs must not be read on another stream, without such an edge. */
void* rb_ctx_get_stream(rb_ctx* ctx);
int rb_ctx_sync(rb_ctx* ctx);
/* Synchronises and returns the sticky status of the asynchronous (device-pointer) calls issued
 * since the last rb_ctx_status(); clears it. */
int rb_ctx_status(rb_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t rb_ctx_launch_count(rb_ctx* ctx);
/* enable != 0 (the default): G2 points passed in by the caller (ciphertext / key members, table bases,
 * pairing arguments) are tested for membership in the order-r subgroup, as rabe_bn does when it
 * decodes a G2 value.  enable == 0 declares them trusted -- e.g. ciphertexts this library produced a
 * moment ago, or typed rabe_bn values that were validated when they were deserialised. */
int rb_ctx_set_g2_subgroup_check(rb_ctx* ctx, int enable);
/* Validates n G2 points (range, on the twist, in the subgroup) regardless of the context setting:
 * RB_OK or RB_ENOTMEMBER.  The decode-time check of rabe_bn's G2, as a batch. */
int rb_g2_check_batch(rb_ctx* ctx, const uint8_t* q, size_t n);
/* Pairing kernels come in two layouts with identical results.  THROUGHPUT: two lanes per Miller loop / final
 * exponentiation -- fewest instructions per product, the best batches/s once several batches are in flight on the
 * GPU.  LATENCY: six lanes per work item, every Fq12 value in registers, a ciphertext's three decrypt terms on one
 * accumulator -- one batch finishes sooner (AC17 decrypt of 4096 items: 7.1 ms instead of 8.4 ms on a B200).
 * AUTO (default): LATENCY unless at least two other contexts of the same GPU still have work of any kind queued or running. */
#define RB_PAIRING_AUTO 0
#define RB_PAIRING_THROUGHPUT 1
#define RB_PAIRING_LATENCY 2
int rb_ctx_set_pairing_layout(rb_ctx* ctx, int mode);
/* enable != 0: calls with HOST buffers enqueue their copies and kernels on the context's stream and return at once;
 * outputs and the status arrive with rb_ctx_sync() / rb_ctx_status() (or an event recorded on rb_ctx_get_stream()).
 * The host buffers of a call must stay alive and untouched until then; use page-locked memory for real overlap.
 * Calls that build a handle or return a verdict (tables, keys, rb_g2_check_batch) still complete before returning.
 * One host thread can then keep every context of a GPU busy.  Default: off (host-buffer calls are synchronous). */
int rb_ctx_set_async(rb_ctx* ctx, int enable);
/* Per-kernel timing with CUDA events on the context's stream: enable, run calls, then read a JSON
 * object {"kernel": {"launches": n, "ms": total}, ...}.  (out == NULL: only *needed is set.) */
int rb_ctx_profile(rb_ctx* ctx, int enable);
int rb_ctx_profile_report(rb_ctx* ctx, char* out, size_t cap, size_t* needed);

/* ---- L0: batched rabe_bn operators ---------------------------------------------------------- */
/* Fq / Fr products, a[i]*b[i] (rabe_bn `Fr * Fr`: ac17/mod.rs:175,208,235; secretsharing:25-28). */
int rb_fq_mul_batch(rb_ctx*, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);
int rb_fr_mul_batch(rb_ctx*, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);
/* Micro-benchmark used for the roofline denominators: every thread performs `iters` dependent
 * Montgomery products; returns nothing but consumes time (results are folded into out[n][32]). */
int rb_fq_mul_chain(rb_ctx*, const uint8_t* a, const uint8_t* b, size_t n, int iters, uint8_t* out);

/* Fixed-base precomputation for `base * k` / `base.pow(k)` with a base that is reused
 * (pk / msk members, generators).  window_bits in [4,26] for G1, [4,16] for G2/Gt.
 * Device memory per table = ceil(256 / w) windows x entries x (64 | 128 | 384) bytes.  G1 tables wider than 12 bits
 * keep SIGNED digits (entries = 2^(w-1) + 32 per window: a negative digit adds the entry with y negated), all others
 * 2^w entries:   G1: w = 16 -> 32 MiB, 20 -> 436 MB, 24 -> 5.9 GB, 26 -> 21.5 GB;  G2: 8 -> 1 MiB, 16 -> 134 MB;
 * Gt: 8 -> 3 MiB, 16 -> 403 MB.  Wider windows trade HBM for mixed additions per output (ceil(256/w) - 1): 15 at 16
 * bits, 10 at 24, 9 at 26.  Tables wider than 12 bits are filled by chunked incremental addition (under a second). */
int rb_g1_table_create(rb_ctx*, const uint8_t base[RB_G1_BYTES], int window_bits, rb_table** out);
int rb_g2_table_create(rb_ctx*, const uint8_t base[RB_G2_BYTES], int window_bits, rb_table** out);
int rb_gt_table_create(rb_ctx*, const uint8_t base[RB_GT_BYTES], int window_bits, rb_table** out);
void rb_table_destroy(rb_table* t);

/* out[i] = base * k[i]        (`G1 * Fr`: ac17:170,235-257,348; bsw:103,147,197,233; lsw:99-152) */
int rb_g1_mul_fixed_batch(rb_ctx*, const rb_table* t, const uint8_t* k, size_t n, uint8_t* out);
/* out[i] = base * k[i]        (`G2 * Fr`: ac17:164,219,300-302; bsw:104-148,198,241; lsw:153,212) */
int rb_g2_mul_fixed_batch(rb_ctx*, const rb_table* t, const uint8_t* k, size_t n, uint8_t* out);
/* out[i] = base.pow(k[i])     (`Gt.pow(Fr)`: ac17:175,359; bsw:234; lsw:103,211; aw11:144,263)   */
int rb_gt_pow_fixed_batch(rb_ctx*, const rb_table* t, const uint8_t* k, size_t n, uint8_t* out);
/* out[i] = p[i] * k[i]        (variable base; bsw:147-148,197-198, lsw decrypt folding)          */
int rb_g1_mul_var_batch(rb_ctx*, const uint8_t* p, const uint8_t* k, size_t n, uint8_t* out);
int rb_g2_mul_var_batch(rb_ctx*, const uint8_t* p, const uint8_t* k, size_t n, uint8_t* out);
/* out[i] = a[i].pow(k[i])     (`Gt.pow(Fr)` with a variable base: bsw:294, lsw:278, aw11:348)   */
int rb_gt_pow_var_batch(rb_ctx*, const uint8_t* a, const uint8_t* k, size_t n, uint8_t* out);
/* out[o] = sum_{j in [offs[o], offs[o+1])} points[idx[j]]   (`G1 + G1` loops: ac17:404-415)      */
int rb_g1_sum_gather_batch(rb_ctx*, const uint8_t* points, size_t n_points, const uint32_t* idx,
                           const uint32_t* offs, size_t n_out, uint8_t* out);
/* out[b] = prod_{j in [offs[b], offs[b+1])} pairing(P[j], Q[j])  -- one final exponentiation per
 * product (`pairing(G1,G2)` and the `Gt * Gt` folds around it: ac17:415-418; bsw:291-293,308;
 * lsw:275-280; aw11:340-350).  An empty range or a pair with an infinite member contributes one. */
int rb_pairing_product_batch(rb_ctx*, const uint8_t* P, const uint8_t* Q, const uint32_t* offs,
                             size_t n_products, uint8_t* out);
/* out[i] = a[i] * b[i] ; out[i] = 1/a[i]   (`Gt * Gt`, `Gt.inverse()`)                           */
int rb_gt_mul_batch(rb_ctx*, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);
int rb_gt_inverse_batch(rb_ctx*, const uint8_t* a, size_t n, uint8_t* out);

/* Fr element-wise operators: op 0 add, 1 sub, 2 mul, 3 inverse(a) (b ignored; zero -> RB_ENOTMEMBER),
 * 4 neg(a).  b_is_scalar != 0 broadcasts b[0] (`Fr` ops: secretsharing/mod.rs:25-28,66,218;
 * bsw/mod.rs:103,147; lsw/mod.rs:94,143-152,204-206). */
int rb_fr_op_batch(rb_ctx*, int op, const uint8_t* a, const uint8_t* b, int b_is_scalar, size_t n, uint8_t* out);
/* out[i] = a[i] + b[i] (b_is_point != 0: + b[0])   (`G + G`: bsw/mod.rs:147-148,197-198; aw11/mod.rs:224,276) */
int rb_g1_add_batch(rb_ctx*, const uint8_t* a, const uint8_t* b, int b_is_point, size_t n, uint8_t* out);
int rb_g2_add_batch(rb_ctx*, const uint8_t* a, const uint8_t* b, int b_is_point, size_t n, uint8_t* out);

/* ---- secret sharing over a policy tree (secretsharing/mod.rs) -------------------------------- */
typedef struct rb_share_plan rb_share_plan;
typedef struct rb_policy rb_policy;
/* Flattens gen_shares_policy (secretsharing/mod.rs:82-141) for one policy and keeps it on the device. */
int rb_share_plan_create(rb_ctx*, const rb_policy*, rb_share_plan** out);
void rb_share_plan_free(rb_share_plan*);
/* n_leaves = shares per item (DFS order, as gen_shares_policy returns them); n_coefs = random
 * coefficients per item = sum over AND gates of (children - 1), in the order rabe draws them. */
int rb_share_plan_dims(const rb_share_plan*, uint32_t* n_leaves, uint32_t* n_coefs);
/* shares[B][n_leaves] from secret[B] and coeffs[B][n_coefs]  (polynomial(), secretsharing:215-221) */
int rb_shares_batch(rb_ctx*, const rb_share_plan*, const uint8_t* secret, const uint8_t* coeffs, size_t B, uint8_t* shares);
/* coeff[n_leaves] = calc_coefficients(policy, Fr::one()) in DFS order (secretsharing/mod.rs:9-72) */
int rb_policy_coefficients(rb_ctx*, const rb_policy*, uint8_t* out);

/* ---- L1: AC17 (FAME) CP-ABE, schemes/ac17/mod.rs ------------------------------------------- */
/* Ac17PublicKey (ac17/mod.rs:62): g[64] | h_a[3][128] | e_gh_ka[2][384] = 1216 bytes.
 * Loading builds the fixed-base tables for g, h_a[0..2] and e_gh_ka[0..1] on the device. */
#define RB_AC17_PK_BYTES 1216
#define RB_AC17_MSK_BYTES 512 /* Ac17MasterKey (ac17/mod.rs:72): g | h[128] | g_k[3][64] | a[2][32] | b[2][32] */
typedef struct rb_ac17_pk rb_ac17_pk;
typedef struct rb_ac17_msk rb_ac17_msk;
typedef struct rb_msp rb_msp;

int rb_ac17_pk_load(rb_ctx*, const uint8_t pk[RB_AC17_PK_BYTES], rb_ac17_pk** out);
/* Same with explicit window widths of the fixed-base tables (pk.g: 4..26 bits, pk.h_a and
 * pk.e_gh_ka: 4..16 bits).  Wider windows trade HBM for work: a 24-bit G1 table is 5.9 GB and
 * cuts the per-output mixed additions of cp_encrypt (ac17/mod.rs:330-356) from 15 to 10; 26 bits
 * is 21.5 GB per key for 9 (see rb_g1_table_create for the sizes).
 * rb_ac17_pk_load == rb_ac17_pk_load_ex(.., 16, 8, 8, ..). */
int rb_ac17_pk_load_ex(rb_ctx*, const uint8_t pk[RB_AC17_PK_BYTES], int g1_window, int g2_window, int gt_window, rb_ac17_pk** out);
void rb_ac17_pk_free(rb_ac17_pk*);
int rb_ac17_msk_load(rb_ctx*, const uint8_t msk[RB_AC17_MSK_BYTES], rb_ac17_msk** out);
void rb_ac17_msk_free(rb_ac17_msk*);

/* setup (ac17/mod.rs:141-188).  rnd = 9 Fr in the order the reference draws them:
 * rho_g, rho_h (g = G1gen*rho_g, h = G2gen*rho_h), a0, b0, a1, b1, k0, k1, k2. */
int rb_ac17_setup(rb_ctx*, const uint8_t rnd[9 * RB_FR_BYTES], uint8_t pk[RB_AC17_PK_BYTES],
                  uint8_t msk[RB_AC17_MSK_BYTES]);

/* A policy in the numeric form the encrypt loops consume (ac17/mod.rs:286-356):
 *   m      n1 x n2 row-major i8 in {-1,0,1}     (AbePolicy.m, msp.rs:11)
 *   h_row  [n1][3][2] Fr = sha3_hash_fr(pi[i] + l + t)          (ac17:333-339 with hash/mod.rs:23)
 *   h_col  [n2][3][2] Fr = sha3_hash_fr("0" + (j+1) + l + t)    (ac17:305-328)
 * The per-row scalar table  A[i][l][t] = h_row[i][l][t] + sum_j m[i][j]*h_col[j][l][t]  is folded
 * on the device when the policy is loaded. */
int rb_msp_load(rb_ctx*, uint32_t n1, uint32_t n2, const int8_t* m, const uint8_t* h_row,
                const uint8_t* h_col, rb_msp** out);
/* n_pol policies of the same shape (n1 rows, n2 columns; pad narrower matrices with zero columns),
 * one per batch item ("distinct" policy mode): m [n_pol][n1][n2], h_row [n_pol][n1][3][2],
 * h_col [n_pol][n2][3][2] (the hashes may come from rb_sha3_fr_batch).  rb_ac17_cp_encrypt_batch
 * with such a handle requires B == n_pol and encrypts item b under policy b. */
int rb_msp_load_batch(rb_ctx*, uint32_t n1, uint32_t n2, const int8_t* m, const uint8_t* h_row,
                      const uint8_t* h_col, size_t n_pol, rb_msp** out);
/* Refolds a handle in place from new m / h_row / h_col of the shape it was loaded with (n_pol, n1,
 * n2 unchanged): no allocation; with device pointers the call is stream-ordered on the context.
 * One handle per context is the intended use when every batch carries new policies (the reference
 * rebuilds these scalars inside every cp_encrypt call, ac17/mod.rs:305-339).
 * h_col_shared != 0: h_col is ONE table [n2][3][2] used by every policy -- the column labels
 * "0"+(j+1)+l+t (ac17:305-328) do not depend on the policy, so a batch needs them hashed once.
 * Matrix entries outside {-1, 0, 1} are detected on the device: RB_EPOLICY from this call (host buffers,
 * synchronous mode) or from the next rb_ctx_status() (device buffers / asynchronous mode). */
int rb_msp_reload_batch(rb_ctx*, rb_msp*, const int8_t* m, const uint8_t* h_row, const uint8_t* h_col,
                        int h_col_shared);
void rb_msp_free(rb_msp*);

/* cp_encrypt, batch of B independent encryptions under one policy (ac17/mod.rs:286-368):
 *   s    [B][2] Fr     the reference's `s` vector (:289-295)
 *   msg  [B] Gt        the reference's random `msg` (:362)
 *   c_0  [B][3] G2     (:297-302)      c  [B][n1][3] G1  (:330-356)      c_p [B] Gt (:357-368) */
int rb_ac17_cp_encrypt_batch(rb_ctx*, const rb_ac17_pk*, const rb_msp*, const uint8_t* s,
                             const uint8_t* msg, size_t B, uint8_t* c_0, uint8_t* c, uint8_t* c_p);

/* cp_keygen, batch of B keys over the same attribute list (ac17/mod.rs:199-261):
 *   h_attr [n][3][2] Fr = sha3_hash_fr(attr + l + t) (:231-235); h_01 [3][2] Fr = sha3_hash_fr("01"+l+t) (:250-254)
 *   rnd    [B][n+3] Fr : r0, r1, sigma_attr[0..n), sigma   (draw order of the reference)
 *   k_0 [B][3] G2 ; k [B][n][3] G1 ; k_p [B][3] G1 */
int rb_ac17_cp_keygen_batch(rb_ctx*, const rb_ac17_msk*, uint32_t n, const uint8_t* h_attr,
                            const uint8_t* h_01, const uint8_t* rnd, size_t B, uint8_t* k_0,
                            uint8_t* k, uint8_t* k_p);

/* cp_decrypt, batch of B ciphertexts with n1 rows each, one secret key (ac17/mod.rs:400-418):
 *   ct_idx / ct_offs   rows of ct.c to add for item b: ct_idx[ct_offs[b] .. ct_offs[b+1])
 *                      (ct_offs == NULL: the n_ct_idx entries of ct_idx apply to every item)
 *   sk_idx / sk_offs   rows of sk.k to add, same convention
 * These lists are what the name-matching loops at :404-413 select (see host layer).
 *   msg_out [B] Gt = c_p * prod_i e(prod_g_i, k_0[i]) / e(k_p[i] + prod_h_i, c_0[i]) */
int rb_ac17_cp_decrypt_batch(rb_ctx*, const uint8_t* k_0, const uint8_t* k, uint32_t n_k,
                             const uint8_t* k_p, const uint8_t* c_0, const uint8_t* c, uint32_t n1,
                             const uint8_t* c_p, size_t B, const uint32_t* ct_idx,
                             const uint32_t* ct_offs, size_t n_ct_idx, const uint32_t* sk_idx,
                             const uint32_t* sk_offs, size_t n_sk_idx, uint8_t* msg_out);

/* kp_keygen, batch of B keys for one policy (ac17/mod.rs:439-546): m / h_row / h_col as in rb_msp_load;
 *   rnd [B][2 + (n2-1) + n1] Fr : r0, r1, sigma'[0..n2-2], sigma_attr[0..n1)   (draw order of the reference)
 *   k_0 [B][3] G2 ; k [B][n1][3] G1   (Ac17KpSecretKey.sk; k_p is empty in the KP variant)
 * kp_encrypt (:556-617) is rb_ac17_cp_encrypt_batch with the degenerate policy (n2 = 1, m = 0,
 * h_row = the attribute hashes); kp_decrypt (:625-680) is rb_ac17_cp_decrypt_batch with k_p = 3 points
 * at infinity (all-zero bytes). */
int rb_ac17_kp_keygen_batch(rb_ctx*, const rb_ac17_msk*, uint32_t n1, uint32_t n2, const int8_t* m, const uint8_t* h_row,
                            const uint8_t* h_col, const uint8_t* rnd, size_t B, uint8_t* k_0, uint8_t* k);

/* A secret key kept on the device (Ac17SecretKey, ac17/mod.rs:113: k_0[3] G2, k[n_k][3] G1, k_p[3] G1).
 * Loading precomputes the Miller-loop lines of k_0[0..2] -- the fixed second arguments of the three
 * pairings `pairing(_prod_g, sk.sk.k_0[_i])` at ac17/mod.rs:416 -- once per key. */
typedef struct rb_ac17_sk rb_ac17_sk;
int rb_ac17_sk_load(rb_ctx*, const uint8_t* k_0, const uint8_t* k, uint32_t n_k, const uint8_t* k_p, rb_ac17_sk** out);
void rb_ac17_sk_free(rb_ac17_sk*);
/* cp_decrypt with a loaded key; arguments as rb_ac17_cp_decrypt_batch. */
int rb_ac17_cp_decrypt_sk_batch(rb_ctx*, const rb_ac17_sk*, const uint8_t* c_0, const uint8_t* c, uint32_t n1,
                                const uint8_t* c_p, size_t B, const uint32_t* ct_idx, const uint32_t* ct_offs,
                                size_t n_ct_idx, const uint32_t* sk_idx, const uint32_t* sk_offs, size_t n_sk_idx,
                                uint8_t* msg_out);

/* ---- host-side policy layer (strings only; no device work) ----------------------------------
 * Mirrors rabe's L2 helpers so that the numeric entry points above can be driven from policy
 * text: utils/policy/pest (parse, serialize_policy), utils/policy/msp.rs (calculate_msp/lw),
 * utils/secretsharing (calc_pruned, node_index), utils/tools (traverse_policy), utils/hash
 * (sha3 -> Fr).  Malformed trees that make rabe panic return RB_EPOLICY. */
#define RB_LANG_JSON 0  /* PolicyLanguage::JsonPolicy  (pest/mod.rs:18) */
#define RB_LANG_HUMAN 1 /* PolicyLanguage::HumanPolicy (pest/mod.rs:20) */

int rb_policy_parse(const char* text, int language, rb_policy** out);                 /* pest/mod.rs:40 */
void rb_policy_free(rb_policy*);
/* writes the serialized policy (NUL terminated) or, when out == NULL, only *needed          */
int rb_policy_serialize(const rb_policy*, int language, char* out, size_t cap, size_t* needed); /* pest/mod.rs:68 */
/* LW matrix (msp.rs:78): *n1 rows, *n2 columns; m (n1*n2, row-major) and names (the n1 row
 * labels, each NUL terminated, concatenated) are filled when non-NULL and large enough.      */
int rb_policy_msp(const rb_policy*, uint32_t* n1, uint32_t* n2, int8_t* m, size_t m_cap, char* names,
                  size_t names_cap, size_t* names_needed);
int rb_policy_satisfied(const rb_policy*, const char* const* attrs, uint32_t n_attrs, int* out); /* tools/mod.rs:31 */
/* calc_pruned (secretsharing/mod.rs:143): out receives n_items pairs "name\0label\0" where
 * label = node_index (name_column).                                                          */
int rb_policy_prune(const rb_policy*, const char* const* attrs, uint32_t n_attrs, int* matched, char* out,
                    size_t cap, size_t* needed, uint32_t* n_items);
int rb_hash_to_fr(const char* s, size_t len, uint8_t out[RB_FR_BYTES]);               /* hash/mod.rs:23 */
/* node_index labels ("name_column", secretsharing/mod.rs:74) of the leaves in DFS order, each NUL
 * terminated -- the order of gen_shares_policy / calc_coefficients results.                      */
int rb_policy_leaf_labels(const rb_policy*, char* out, size_t cap, size_t* needed, uint32_t* n_leaves);

/* AC17 glue: hashes the row labels / column indices as ac17/mod.rs:305-339 does and loads the
 * folded policy onto the device (see rb_msp_load). */
int rb_ac17_msp_from_policy(rb_ctx*, const rb_policy*, rb_msp** out);
/* h_attr [n][3][2] Fr and h_01 [3][2] Fr for rb_ac17_cp_keygen_batch (ac17/mod.rs:231-235,250-254) */
int rb_ac17_attr_hashes(const char* const* attrs, uint32_t n_attrs, uint8_t* h_attr, uint8_t h_01[6 * RB_FR_BYTES]);
/* The gather lists of cp_decrypt: for every entry of calc_pruned(sk_attrs, policy) every ct row /
 * key row with that name (ac17/mod.rs:404-413).  *matched == 0 reproduces rabe's
 * "attributes in sk do not match policy in ct" error (:389,:425). */
int rb_ac17_decrypt_lists(const rb_policy*, const char* const* sk_attrs, uint32_t n_sk, const char* const* ct_names,
                          uint32_t n_ct, int* matched, uint32_t* ct_idx, size_t ct_cap, uint32_t* n_ct_idx,
                          uint32_t* sk_idx, size_t sk_cap, uint32_t* n_sk_idx);

/* SHA3-256 -> Fr for n byte strings on the device (utils/hash/mod.rs:23-31, the scalar behind every
 * sha3_hash(g, s) = g * H(s)): string i = data[offs[i] .. offs[i+1]); out [n] canonical Fr.
 * rb_hash_to_fr is the host-side single-string twin. */
int rb_sha3_fr_batch(rb_ctx*, const uint8_t* data, const uint32_t* offs, size_t n, uint8_t* out);
/* Same, with the total length of data (= offs[n]) stated by the caller, so that device-resident
 * offs need no device->host read and the call stays asynchronous on the context's stream. */
int rb_sha3_fr_batch_len(rb_ctx*, const uint8_t* data, size_t data_len, const uint32_t* offs, size_t n, uint8_t* out);

/* ---- fused batch entry points of BSW / LSW / AW11 -------------------------------------------
 * Same conventions as the AC17 entry points: B independent items per call, every buffer host or
 * device, all randomness explicit, intermediates never leave the device.  leaf_hash[i] =
 * rb_hash_to_fr(attribute of leaf i) in the DFS leaf order of rb_policy_leaf_labels.          */
typedef struct rb_bsw_pk rb_bsw_pk;
/* fixed-base tables of bsw::CpAbePublicKey{g1, g2, h, e_gg_alpha} (bsw/mod.rs:43)             */
int rb_bsw_pk_load(rb_ctx*, const uint8_t g1[RB_G1_BYTES], const uint8_t g2[RB_G2_BYTES], const uint8_t h[RB_G1_BYTES],
                   const uint8_t e_gg_alpha[RB_GT_BYTES], rb_bsw_pk** out);
void rb_bsw_pk_free(rb_bsw_pk*);
/* bsw::encrypt (bsw/mod.rs:217-251): secret [B], coeffs [B][n_coefs] (gen_shares draws, plan order),
 * msg [B] Gt -> c [B] G1, c_p [B] Gt, cy_g1 [B][n] G1, cy_g2 [B][n] G2 (n = leaves of the plan).   */
int rb_bsw_encrypt_batch(rb_ctx*, const rb_bsw_pk*, const rb_share_plan*, const uint8_t* leaf_hash, const uint8_t* secret,
                         const uint8_t* coeffs, const uint8_t* msg, size_t B, uint8_t* c, uint8_t* c_p, uint8_t* cy_g1,
                         uint8_t* cy_g2);
/* bsw::keygen (bsw/mod.rs:125-152): r [B], r_j [B][n] -> d [B] G2, dj_g1 [B][n] G1, dj_g2 [B][n] G2.
 * n == 0 (rabe returns None) is RB_EINVAL.                                                        */
int rb_bsw_keygen_batch(rb_ctx*, const rb_bsw_pk*, const uint8_t beta[RB_FR_BYTES], const uint8_t g2_alpha[RB_G2_BYTES],
                        const uint8_t* attr_hash, uint32_t n, const uint8_t* r, const uint8_t* r_j, size_t B, uint8_t* d,
                        uint8_t* dj_g1, uint8_t* dj_g2);
/* bsw::delegate (bsw/mod.rs:162-206): B delegated keys over one attribute subset of size n; dj_g1 /
 * dj_g2 [n] = the source key's members in subset order, f = pk.f; r [B], r_j [B][n] ->
 * d [B] G2, g1 [B][n] G1, g2 [B][n] G2.                                                            */
int rb_bsw_delegate_batch(rb_ctx*, const rb_bsw_pk*, const uint8_t f[RB_G2_BYTES], const uint8_t d[RB_G2_BYTES], const uint8_t* dj_g1,
                          const uint8_t* dj_g2, const uint8_t* attr_hash, uint32_t n, const uint8_t* r, const uint8_t* r_j, size_t B,
                          uint8_t* d_out, uint8_t* g1_out, uint8_t* g2_out);
/* bsw::decrypt up to the KEM (bsw/mod.rs:260-308): one key, B ciphertexts of one policy.  ct_idx /
 * sk_idx [nI]: positions of the pruned leaves (calc_pruned) in c_y / d_j; coeff [nI]: their
 * calc_coefficients values.  2 nI + 1 Miller loops and ONE final exponentiation per item instead
 * of 2 nI + 1 full pairings and nI Gt exponentiations; out [B] = the Gt value `_msg`.            */
int rb_bsw_decrypt_batch(rb_ctx*, const uint8_t d[RB_G2_BYTES], const uint8_t* dj_g1, const uint8_t* dj_g2, uint32_t n_k,
                         const uint8_t* c, const uint8_t* c_p, const uint8_t* cy_g1, const uint8_t* cy_g2, uint32_t n,
                         const uint32_t* ct_idx, const uint32_t* sk_idx, const uint8_t* coeff, uint32_t nI, size_t B,
                         uint8_t* out);
/* lsw::keygen, positive leaves (lsw/mod.rs:121-160): g1_tab / g2_tab = tables of pk.g1 / pk.g2;
 * coeffs [B][n_coefs], rnd [B][n] -> d1 [B][n] G1, d2 [B][n] G2.                                  */
int rb_lsw_keygen_batch(rb_ctx*, const rb_table* g1_tab, const rb_table* g2_tab, const rb_share_plan*, const uint8_t* leaf_hash,
                        const uint8_t alpha1[RB_FR_BYTES], const uint8_t alpha2[RB_FR_BYTES], const uint8_t* coeffs,
                        const uint8_t* rnd, size_t B, uint8_t* d1, uint8_t* d2);
/* lsw::decrypt up to the KEM (lsw/mod.rs:228-280): sk_d1 / sk_d2 [n_k] = dj members 1 and 2; e1 [B]
 * Gt, e2 [B] G2, ej1 [B][n] = member 1 of every ej tuple; lists as for rb_bsw_decrypt_batch.       */
int rb_lsw_decrypt_batch(rb_ctx*, const uint8_t* sk_d1, const uint8_t* sk_d2, uint32_t n_k, const uint8_t* e1, const uint8_t* e2,
                         const uint8_t* ej1, uint32_t n, const uint32_t* ct_idx, const uint32_t* sk_idx, const uint8_t* coeff,
                         uint32_t nI, size_t B, uint8_t* out);
/* lsw::encrypt (lsw/mod.rs:180-219) for B messages over one attribute list.  rb_lsw_pk holds the
 * fixed-base tables of KpAbePublicKey{g1, g2, g1_b, g1_b2, h_b, e_gg_alpha} (lsw/mod.rs:44).
 * attr_hash [n] = rb_hash_to_fr(attribute i); secret [B]; draws [B][n] = the n scalars the reference
 * draws after `secret` (pushed as sx[1..n], :197-200); msg [B] Gt ->
 * e1 [B] Gt, e2 [B] G2, ej1 / ej2 / ej3 [B][n] G1 = members 1..3 of the ej tuples.  The reference's
 * `sx[0] = sx[0] - sx[_i]` quirk (sx[0] ends as minus the sum of sx[1..n-1]) is reproduced.        */
typedef struct rb_lsw_pk rb_lsw_pk;
int rb_lsw_pk_load(rb_ctx*, const uint8_t g1[RB_G1_BYTES], const uint8_t g2[RB_G2_BYTES], const uint8_t g1_b[RB_G1_BYTES],
                   const uint8_t g1_b2[RB_G1_BYTES], const uint8_t h_b[RB_G1_BYTES], const uint8_t e_gg_alpha[RB_GT_BYTES],
                   rb_lsw_pk** out);
void rb_lsw_pk_free(rb_lsw_pk*);
int rb_lsw_encrypt_batch(rb_ctx*, const rb_lsw_pk*, const uint8_t* attr_hash, uint32_t n, const uint8_t* secret, const uint8_t* draws,
                         const uint8_t* msg, size_t B, uint8_t* e1, uint8_t* e2, uint8_t* ej1, uint8_t* ej2, uint8_t* ej3);
/* ghw11::transform (ghw11/mod.rs:227-294), the outsourced half of GHW11 decryption: one transform
 * key {k_z, l_z, kx[n_k]}, B ciphertexts of one policy with c1 [B] G1 and the ci_di members
 * ci / di [B][n] G1; ct_idx / sk_idx / coeff [nI] as for rb_bsw_decrypt_batch -> t [B] Gt
 * (Ghw11TransformCiphertext.t; .c is the ciphertext's c unchanged).  All 2 nI + 1 pairings have
 * key-side G2 arguments: line tables once per call, ONE final exponentiation per item.            */
int rb_ghw11_transform_batch(rb_ctx*, const uint8_t k_z[RB_G2_BYTES], const uint8_t l_z[RB_G2_BYTES], const uint8_t* kx, uint32_t n_k,
                             const uint8_t* c1, const uint8_t* ci, const uint8_t* di, uint32_t n, const uint32_t* ct_idx,
                             const uint32_t* sk_idx, const uint8_t* coeff, uint32_t nI, size_t B, uint8_t* t);
/* ghw11::decrypt_out up to the KEM (ghw11/mod.rs:297-305): msg[b] = c[b] * (t[b]^z)^-1.           */
int rb_ghw11_decrypt_out_batch(rb_ctx*, const uint8_t* c, const uint8_t* t, const uint8_t z[RB_FR_BYTES], size_t B, uint8_t* msg);
/* aw11::encrypt (aw11/mod.rs:241-289) when every leaf has an authority key: g2_tab / egg_tab =
 * tables of gk.g2 / e(g1,g2); pk_gt [n] / pk_g2 [n] = the (Gt, G2) members of each leaf's
 * Aw11PublicKey entry; s [B], s_coeffs / w_coeffs [B][n_coefs], r_x [B][n], msg [B] ->
 * c_0 [B] Gt, c1 [B][n] Gt, c2 / c3 [B][n] G2.  The constant pairing(g1,g2) that the reference
 * recomputes per row (:274) is the table base.                                                    */
int rb_aw11_encrypt_batch(rb_ctx*, const rb_table* g2_tab, const rb_table* egg_tab, const rb_share_plan*, const uint8_t* pk_gt,
                          const uint8_t* pk_g2, const uint8_t* s, const uint8_t* s_coeffs, const uint8_t* w_coeffs,
                          const uint8_t* r_x, const uint8_t* msg, size_t B, uint8_t* c_0, uint8_t* c1, uint8_t* c2, uint8_t* c3);

/* aw11::decrypt up to the KEM (aw11/mod.rs:298-350): one user key, B ciphertexts of one policy.
 * h = sha3_hash(g1, gid) (:317), sk_k [n_k] = the G1 members of sk.attr; c_0 [B], c1 [B][n] Gt,
 * c2 / c3 [B][n] G2; ct_idx / sk_idx / coeff [nI] as for rb_bsw_decrypt_batch.  out [B] = msg.  */
int rb_aw11_decrypt_batch(rb_ctx*, const uint8_t h[RB_G1_BYTES], const uint8_t* sk_k, uint32_t n_k, const uint8_t* c_0,
                          const uint8_t* c1, const uint8_t* c2, const uint8_t* c3, uint32_t n, const uint32_t* ct_idx,
                          const uint32_t* sk_idx, const uint8_t* coeff, uint32_t nI, size_t B, uint8_t* out);
/* Fixed-base tables of the (Gt, G2) members of n authority-key attributes (Aw11PublicKey.attr,
 * aw11/mod.rs:59) and aw11::encrypt over them: pk_attr.1^r_x and pk_attr.2 * r_x (:274,:276)
 * become table walks.  leaf_attr [n_leaves]: attribute index of every policy leaf (NULL: identity). */
typedef struct rb_aw11_pk rb_aw11_pk;
int rb_aw11_pk_load(rb_ctx*, const uint8_t* pk_gt, const uint8_t* pk_g2, uint32_t n, rb_aw11_pk** out);
void rb_aw11_pk_free(rb_aw11_pk*);
int rb_aw11_encrypt_pk_batch(rb_ctx*, const rb_table* g2_tab, const rb_table* egg_tab, const rb_share_plan*, const rb_aw11_pk*,
                             const uint32_t* leaf_attr, const uint8_t* s, const uint8_t* s_coeffs, const uint8_t* w_coeffs,
                             const uint8_t* r_x, const uint8_t* msg, size_t B, uint8_t* c_0, uint8_t* c1, uint8_t* c2, uint8_t* c3);

/* ---- KEM tail: rabe's encrypt_symmetric / decrypt_symmetric for a batch (utils/aes/mod.rs:10-55) ---------------
 * key[b] = SHA3-256(canonical Gt bytes of gt[b]) (`kdf`, :47-55); AES-256-GCM, 12-byte nonce, no associated data.
 * encrypt: item b = data[offs[b] .. offs[b+1]); out receives  nonce | ciphertext | 16-byte tag  per item, item b at
 *          offs[b] + 28 b (total offs[B] + 28 B bytes) -- the `[nonce|ciphertext]` layout of :21.  nonce [B][12] is an
 *          explicit input like all randomness (the reference draws it from thread_rng, :17).
 * decrypt: item b = nonce_ct[offs[b] .. offs[b+1]) in that layout; out (offs[B] bytes) receives plaintext b -- 28 bytes
 *          shorter than its blob -- AT offs[b]; ok[b] = 1 iff the tag verifies (rabe: `decryption error`, :41-44), else 0
 *          and the item's output is zeroed (a blob shorter than 28 bytes is a forgery).
 * offs [B+1] must be a host array (it sizes the copies); gt / nonce / data / out may be host or device buffers. */
int rb_kem_encrypt_batch(rb_ctx*, const uint8_t* gt, const uint8_t* nonce, const uint8_t* data, const uint32_t* offs, size_t B, uint8_t* out);
int rb_kem_decrypt_batch(rb_ctx*, const uint8_t* gt, const uint8_t* nonce_ct, const uint32_t* offs, size_t B, uint8_t* out, int* ok);

#ifdef __cplusplus
}
#endif
#endif /* RABE_B200_H */
