// TEST INFRASTRUCTURE ONLY: the six-lane Fq12 layer of rabe_b200/csrc/wide.cuh compiled for the HOST
// (-DRB_HOST_SIM).  A lane is a host thread, a warp shuffle is a barrier-protected exchange through a shared
// buffer, so the SAME source that runs on the device (minus the PTX carry chains, which have portable twins) is
// compared here with the one-thread tower / pairing code of tower.cuh / pairing.cuh.  Never linked into the product.
#define RB_HOST_SIM 1
#include "../../rabe_b200/csrc/wide.cuh"
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>
using namespace rb;

namespace rb { namespace w6 {
struct Team {
  Fp buf[LANES]; uint32_t wbuf[LANES];
  std::atomic<int> count{0}; std::atomic<int> gen{0};
  void barrier() {
    int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) == LANES - 1) { count.store(0, std::memory_order_relaxed); gen.store(g + 1, std::memory_order_release); }
    else while (gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
  }
};
Fp team_exchange(const Lane& L, const Fp& mine, int src) {
  L.team->buf[L.k] = mine; L.team->barrier();
  Fp r = L.team->buf[((src % LANES) + LANES) % LANES]; L.team->barrier();
  return r;
}
uint32_t team_exchange_u32(const Lane& L, uint32_t mine, int src) {
  L.team->wbuf[L.k] = mine; L.team->barrier();
  uint32_t r = L.team->wbuf[((src % LANES) + LANES) % LANES]; L.team->barrier();
  return r;
}
}}
using namespace rb::w6;

template <class F> static void run_team(F fn) {
  Team team;
  std::vector<std::thread> th;
  for (int k = 0; k < LANES; ++k) th.emplace_back([&, k] { Lane L{k, &team}; fn(L); });
  for (auto& t : th) t.join();
}
static Fp2 coeff(const Fp12& x, int k) { return f12c(x, tower_index(k)); }
static void set_coeff(Fp12& x, int k, const Fp2& v) { f12c(x, tower_index(k)) = v; }

extern "C" {
// op: 0 mul, 1 sqr, 2 cyclotomic_sqr, 3 inverse, 4 frobenius(j = arg), 5 conj, 6 final_exponentiation, 7 mul_line (b = l0|l3|l4 in the first 192 bytes)
void ws_fp12_op(int op, int arg, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fp12 x, y, r; fp12_load_be(x, a); if (b) fp12_load_be(y, b); else y = x;
  run_team([&](Lane L) {
    Fp2 f = coeff(x, L.k), g = coeff(y, L.k), o;
    switch (op) {
      case 0: o = mul(L, f, g); break;
      case 1: o = sqr(L, f); break;
      case 2: o = cyclotomic_sqr(L, f); break;
      case 3: o = inverse(L, f); break;
      case 4: o = frobenius(L, f, arg); break;
      case 5: o = conj(L, f); break;
      case 6: o = final_exponentiation(L, f); break;
      default: o = mul_line(L, f, f12c(y, 0), f12c(y, 1), f12c(y, 2)); break;
    }
    set_coeff(r, L.k, o);                     // distinct lanes write distinct coefficients
  });
  fp12_store_be(out, r);
}
// the same operations through the one-thread tower (reference for the comparison)
void ws_fp12_ref(int op, int arg, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fp12 x, y, r; fp12_load_be(x, a); if (b) fp12_load_be(y, b); else y = x;
  switch (op) {
    case 0: fp12_mul_to(&r, &x, &y); break;
    case 1: fp12_sqr_to(&r, &x); break;
    case 2: fp12_cyclotomic_sqr_to(&r, &x); break;
    case 3: fp12_inv_to(&r, &x); break;
    case 4: fp12_frobenius_to(&r, &x, arg); break;
    case 5: fp12_conj_to(&r, &x); break;
    case 6: final_exponentiation(&r, &x); break;
    default: { Fp2 l0 = f12c(y, 0), l3 = f12c(y, 1), l4 = f12c(y, 2); r = x; fp12_mul_by_line(&r, &l0, &l3, &l4); break; }
  }
  fp12_store_be(out, r);
}
// sum_t x_t * y_t / R mod N through the wide accumulator (n <= 6 products of Montgomery values; flags bit t: use x_t + x_t unreduced)
void ws_dot(const uint8_t* xs, const uint8_t* ys, int n, uint8_t* out) {
  WAcc A; wacc_zero(A);
  for (int t = 0; t < n; ++t) {
    Fp x = fe_to_mont(fe_load_be<ModP>(xs + 32 * t)), y = fe_to_mont(fe_load_be<ModP>(ys + 32 * t));
    wacc_mac(A, add_nr(x, x), add_nr(y, y));             // factors up to 2N - 2: the largest the layer ever feeds
  }
  fe_store_be(out, fe_from_mont(wacc_redc<24>(A)));
}
// Miller product of up to three terms on one accumulator, then the final exponentiation.
//   pv, pf: [3][64] canonical G1; q, qf: [3][128] canonical G2; mask bit 2j: term j has (pv, q), bit 2j+1: it has (pf, qf)
// returns FE(prod_j miller(pv_j, q_j) miller(pf_j, qf_j)), to be compared with the product of the pairings
void ws_pairing_terms(const uint8_t* pv, const uint8_t* q, const uint8_t* pf, const uint8_t* qf, int mask, int unit, uint8_t* out) {
  static MillerLine lines[3][MILLER_LINES], raw[MILLER_LINES];
  G1Affine gen1; gen1.x = fe_one<ModP>(); gen1.y = fe_dbl(fe_one<ModP>());
  G2Affine gen2; gen2.x = G2_GEN_X; gen2.y = G2_GEN_Y;
  G1Affine PV[3], PF[3]; G2Affine Q[3];
  for (int j = 0; j < 3; ++j) {
    const bool hv = (mask >> (2 * j)) & 1, hf = (mask >> (2 * j + 1)) & 1;
    PV[j] = hv ? g1_load_be(pv + 64 * j) : gen1; Q[j] = hv ? g2_load_be(q + 128 * j) : gen2;
    PF[j] = hf ? g1_load_be(pf + 64 * j) : gen1;
    G2Affine QF = hf ? g2_load_be(qf + 128 * j) : gen2;
    if (unit) {                                     // the table of a loaded key: every line divided by its l0
      miller_lines_for(raw, &QF);
      if (!miller_lines_normalize(lines[j], raw, MILLER_LINES)) { memset(out, 0xff, 384); return; }
    } else miller_lines_for(lines[j], &QF);
  }
  Fp12 r;
  run_team([&](Lane L) {
    const int j = L.k / 2;
    PairState s;
    s.t.x = Q[j].x; s.t.y = Q[j].y; s.t.z = fp2_one(); s.qx = Q[j].x; s.qy = Q[j].y;
    s.xv = PV[j].x; s.yv = PV[j].y; s.xf = PF[j].x; s.yf = PF[j].y; s.lines = lines[j];
    s.has_v = (mask >> (2 * j)) & 1; s.has_f = (mask >> (2 * j + 1)) & 1; s.unit_fixed = unit != 0;
    Fp2 f = miller_terms(L, &s, 3);
    f = final_exponentiation(L, f);
    set_coeff(r, L.k, f);
  });
  fp12_store_be(out, r);
}
unsigned long long ws_mul_count() { return g_host_mul_count; }
void ws_mul_count_reset() { g_host_mul_count = 0; }
}
