// Optimal-ate pairing on BN254 for sm_100a: Miller loop over 6u+2 with the G2 argument kept in
// homogeneous projective coordinates on the twist, sparse line multiplication, and the final
// exponentiation f^((p^6-1)(p^2+1)) followed by the hard part f^LAMBDA with
//   LAMBDA = p^3(12u^3+6u^2+4u-1) + p^2(12u^3+6u^2+6u) + p(12u^3+6u^2+4u) + (12u^3+12u^2+6u+1)
// = 2u(6u^2+3u+1) * (p^4-p^2+1)/r  -- the exponent of the zcash-bn lineage that rabe-bn 0.4.23
// forks, so Gt values agree with the oracle bit for bit.
//
// Replaces `rabe_bn::pairing(G1, G2)` at /root/reference/src/schemes/ac17/mod.rs:148,415-416,
// bsw/mod.rs:108,292-293,308, lsw/mod.rs:103,275-276, aw11/mod.rs:144,263,274,340-341.
// A product of pairings shares one final exponentiation (the Miller values are multiplied first).
#pragma once
#include "curve.cuh"

namespace rb {

struct MillerLine { Fp2 l0, l3, l4; };        // line value = l0 + (l3*yP) w^3 + (l4*xP) w^4
typedef MillerLine FullLine;
constexpr int MILLER_LINES = (ATE_NAF_LEN - 1) + 21 + 2;   // 65 tangents + 21 chords + 2 Frobenius chords


#include "pairing_body.inc"

}  // namespace rb
