// TEST INFRASTRUCTURE ONLY -- see bn254.hpp for provenance and the "parity unpinned" note.
#include "bn254.hpp"

namespace orc {

thread_local OpCount g_ops;

static U256 u256_div_small(const U256& a, u64 d) {
  U256 q{}; u128 rem = 0;
  for (int i = 3; i >= 0; --i) { u128 cur = (rem << 64) | a.v[i]; q.v[i] = (u64)(cur / d); rem = cur % d; }
  return q;
}

const FrobConsts& frob() {
  static const FrobConsts F = [] {
    FrobConsts f;
    Fq2 xi = {Fq::from_u64(9), Fq::one()};
    U256 one{{1, 0, 0, 0}};
    U256 e = u256_div_small(u256_sub(MOD_P, one), 6);         // (p-1)/6
    Fq2 g = xi.pow(e);                                        // xi^((p-1)/6)
    Fq2 n = g * g.conj();                                     // xi^((p^2-1)/6), lies in Fq
    Fq2 h = n * g;                                            // xi^((p^3-1)/6)
    f.g1[0] = f.g2[0] = f.g3[0] = Fq2::one();
    for (int k = 1; k < 6; ++k) { f.g1[k] = f.g1[k - 1] * g; f.g2[k] = f.g2[k - 1] * n; f.g3[k] = f.g3[k - 1] * h; }
    f.tw_x1 = f.g1[2]; f.tw_y1 = f.g1[3]; f.tw_x2 = f.g2[2]; f.tw_y2 = f.g2[3];
    Fq2 three = {Fq::from_u64(3), Fq::zero()};
    f.twist_b = three * xi.inverse();
    f.two_inv = Fq::from_u64(2).inverse();
    return f;
  }();
  return F;
}

G1 g1_generator() { return G1::from_affine(Fq::from_u64(1), Fq::from_u64(2)); }

static Fq fq_from_dec(const char* s) {
  Fq acc = Fq::zero(), ten = Fq::from_u64(10);
  for (; *s; ++s) acc = acc * ten + Fq::from_u64((u64)(*s - '0'));
  return acc;
}

G2 g2_generator() {
  static const G2 g = [] {
    Fq2 x = {fq_from_dec("10857046999023057135944570762232829481370756359578518086990519993285655852781"),
             fq_from_dec("11559732032986387107991004021392285783925812861821192530917403151452391805634")};
    Fq2 y = {fq_from_dec("8495653923123431417604973247489272438418190587263600148770280649306958101930"),
             fq_from_dec("4082367875863433681332203403145435568316851327593401208105741076214120093531")};
    return G2::from_affine(x, y);
  }();
  return g;
}

void g1_to_bytes(const G1& p, uint8_t out[64]) {
  Fq x, y;
  if (!p.to_affine(x, y)) { memset(out, 0, 64); return; }
  x.to_be(out); y.to_be(out + 32);
}
bool g1_from_bytes(const uint8_t in[64], G1& p) {
  bool allz = true; for (int i = 0; i < 64; ++i) allz &= in[i] == 0;
  if (allz) { p = G1::zero(); return true; }
  Fq x, y;
  if (!Fq::from_be(in, x) || !Fq::from_be(in + 32, y)) return false;
  if (y.sqr() != x.sqr() * x + Fq::from_u64(3)) return false;
  p = G1::from_affine(x, y); return true;
}
void g2_to_bytes(const G2& p, uint8_t out[128]) {
  Fq2 x, y;
  if (!p.to_affine(x, y)) { memset(out, 0, 128); return; }
  x.a.to_be(out); x.b.to_be(out + 32); y.a.to_be(out + 64); y.b.to_be(out + 96);
}
bool g2_from_bytes(const uint8_t in[128], G2& p) {
  bool allz = true; for (int i = 0; i < 128; ++i) allz &= in[i] == 0;
  if (allz) { p = G2::zero(); return true; }
  Fq2 x, y;
  if (!Fq::from_be(in, x.a) || !Fq::from_be(in + 32, x.b) || !Fq::from_be(in + 64, y.a) || !Fq::from_be(in + 96, y.b)) return false;
  if (y.sqr() != x.sqr() * x + frob().twist_b) return false;
  p = G2::from_affine(x, y); return true;
}

// ---------------------------------------------------------------------------------------------
// Miller loop.  T is kept in homogeneous projective coordinates on the twist; every line value is
// the untwisted chord/tangent scaled by an Fq2 factor and by w^3 (both vanish under the final
// exponentiation):  l = l0 + (l3 * yP) w^3 + (l4 * xP) w^4.
struct Homog { Fq2 x, y, z; };

static void dbl_step(Homog& t, Fq2& l0, Fq2& l3, Fq2& l4) {
  const FrobConsts& F = frob();
  Fq2 a = (t.x * t.y).scale(F.two_inv);
  Fq2 b = t.y.sqr();
  Fq2 c = t.z.sqr();
  Fq2 e = F.twist_b * (c.dbl() + c);
  Fq2 f = e.dbl() + e;
  Fq2 g = (b + f).scale(F.two_inv);
  Fq2 h = (t.y + t.z).sqr() - (b + c);
  Fq2 i = e - b;
  Fq2 j = t.x.sqr();
  Fq2 e2 = e.sqr();
  t.x = a * (b - f);
  t.y = g.sqr() - (e2.dbl() + e2);
  t.z = b * h;
  l0 = i.mul_xi();
  l3 = h.neg();
  l4 = j.dbl() + j;
}

static void add_step(Homog& t, const Fq2& qx, const Fq2& qy, Fq2& l0, Fq2& l3, Fq2& l4) {
  Fq2 d = t.x - qx * t.z;
  Fq2 e = t.y - qy * t.z;
  Fq2 f = d.sqr();
  Fq2 g = e.sqr();
  Fq2 h = d * f;
  Fq2 i = t.x * f;
  Fq2 j = h + t.z * g - i.dbl();
  t.x = d * j;
  t.y = e * (i - j) - h * t.y;
  t.z = t.z * h;
  l0 = (e * qx - d * qy).mul_xi();
  l4 = e.neg();
  l3 = d;
}

Fq12 miller_loop(const G1& p, const G2& q) {
  Fq px, py; Fq2 qx, qy;
  if (!p.to_affine(px, py) || !q.to_affine(qx, qy)) return Fq12::one();
  const FrobConsts& F = frob();
  Homog t = {qx, qy, Fq2::one()};
  Fq12 f = Fq12::one();
  Fq2 l0, l3, l4;
  // 6u+2 = 0x1_9d797039be763ba8 (65 bits); the top bit is consumed by T = Q.
  const u64 lo = 0x9d797039be763ba8ull;
  for (int i = 63; i >= 0; --i) {
    f = f.sqr();
    dbl_step(t, l0, l3, l4);
    f = f.mul_by_line(l0, l3.scale(py), l4.scale(px));
    if ((lo >> i) & 1) {
      add_step(t, qx, qy, l0, l3, l4);
      f = f.mul_by_line(l0, l3.scale(py), l4.scale(px));
    }
  }
  Fq2 q1x = qx.conj() * F.tw_x1, q1y = qy.conj() * F.tw_y1;           // pi(Q)
  Fq2 q2x = qx * F.tw_x2, q2y = (qy * F.tw_y2).neg();                  // -pi^2(Q)
  add_step(t, q1x, q1y, l0, l3, l4);
  f = f.mul_by_line(l0, l3.scale(py), l4.scale(px));
  add_step(t, q2x, q2y, l0, l3, l4);
  f = f.mul_by_line(l0, l3.scale(py), l4.scale(px));
  return f;
}

Fq12 final_exponentiation(const Fq12& f) {
  // easy part: f^((p^6-1)(p^2+1))
  Fq12 a = f.conj() * f.inverse();
  Fq12 x = a.frobenius(2) * a;
  // hard part: x^LAMBDA by the lineage's addition chain (exp by -u = conj(x^u))
  auto exp_neg_u = [](const Fq12& v) { return v.cyclotomic_exp_u().conj(); };
  Fq12 A = exp_neg_u(x);
  Fq12 B = A.cyclotomic_sqr();
  Fq12 C = B.cyclotomic_sqr();
  Fq12 D = C * B;
  Fq12 E = exp_neg_u(D);
  Fq12 Fv = E.cyclotomic_sqr();
  Fq12 G = exp_neg_u(Fv);
  Fq12 H = D.conj();
  Fq12 I = G.conj();
  Fq12 J = I * E;
  Fq12 K = J * H;
  Fq12 L = K * B;
  Fq12 M = K * E;
  Fq12 N = M * x;
  Fq12 O = L.frobenius(1);
  Fq12 Pv = O * N;
  Fq12 Q = K.frobenius(2);
  Fq12 Rv = Q * Pv;
  Fq12 S = x.conj();
  Fq12 T = S * L;
  Fq12 Uv = T.frobenius(3);
  return Uv * Rv;
}

Fq12 pairing(const G1& p, const G2& q) {
  if (p.is_zero() || q.is_zero()) return Fq12::one();
  return final_exponentiation(miller_loop(p, q));
}

// ---------------------------------------------------------------------------------------------
// SHA3-256
static inline u64 rotl64(u64 x, int s) { return s ? (x << s) | (x >> (64 - s)) : x; }

static void keccak_f(u64 st[25]) {
  static const u64 RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int round = 0; round < 24; ++round) {
    u64 c[5], d[5], b[25];
    for (int x = 0; x < 5; ++x) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; ++i) st[i] ^= d[i % 5];
    for (int x = 0; x < 5; ++x)
      for (int y = 0; y < 5; ++y) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x + 5 * y]);
    for (int y = 0; y < 5; ++y)
      for (int x = 0; x < 5; ++x) st[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[round];
  }
}

void sha3_256(const uint8_t* data, size_t len, uint8_t out[32]) {
  const size_t rate = 136;
  u64 st[25]; memset(st, 0, sizeof st);
  uint8_t block[136];
  while (len >= rate) {
    for (size_t i = 0; i < rate / 8; ++i) { u64 w; memcpy(&w, data + 8 * i, 8); st[i] ^= w; }
    keccak_f(st); data += rate; len -= rate;
  }
  memset(block, 0, rate);
  memcpy(block, data, len);
  block[len] ^= 0x06; block[rate - 1] ^= 0x80;
  for (size_t i = 0; i < rate / 8; ++i) { u64 w; memcpy(&w, block + 8 * i, 8); st[i] ^= w; }
  keccak_f(st);
  memcpy(out, st, 32);
}

Fr sha3_hash_fr(const std::string& s) {
  uint8_t d[32];
  sha3_256((const uint8_t*)s.data(), s.size(), d);
  return Fr::from_be_reduce(d);
}

}  // namespace orc
