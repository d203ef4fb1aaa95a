#!/usr/bin/env python3
"""Writes tests/golden/bn254_kat.json and tests/golden/ac17_config1.json.

bn254_kat.json comes from oracle/pyref.py ALONE -- the independent pure-Python statement of BN254
(flat Fp12, affine textbook Miller loop, one big-exponent final power): scalar multiples of the
generators, lineage pairings and a Gt power, as canonical big-endian hex.  One entry is a public
known answer that does not come from this repository: 2*G1 on alt_bn128 (the EIP-196 ecMul vector).
ac17_config1.json is BASELINE.json config 1 (4-attribute AND policy, one round trip) run through
the C++ oracle with seeded randomness: SHA-256 digests of every key / ciphertext member plus the
decrypted Gt value.  tests/test_golden.py checks the C++ oracle (CPU) and the CUDA path (GPU)
against both files, so neither needs pyref's minutes of big-integer arithmetic at test time."""
import hashlib, json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from oracle import pyref as r
from oracle import policy as opol

h32 = lambda x: "%064x" % x
g1h = lambda p: "00" * 64 if p is None else h32(p[0]) + h32(p[1])
g2h = lambda q: "00" * 128 if q is None else "".join(h32(x) for x in (q[0][0], q[0][1], q[1][0], q[1][1]))
gth = lambda t: "".join(h32(x) for x in t)

G1 = (1, 2)
G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781, 11559732032986387107991004021392285783925812861821192530917403151452391805634),
      (8495653923123431417604973247489272438418190587263600148770280649306958101930, 4082367875863433681332203403145435568316851327593401208105741076214120093531))
rng = random.Random(20261017)
scalars = [1, 2, 3, r.R - 1, 0xdeadbeef, 1 << 253] + [rng.randrange(r.R) for _ in range(6)]
kat = {"source": "oracle/pyref.py (pure Python, independent of the C++ oracle and of the CUDA code)",
       "g1_mul": [{"k": h32(k), "out": g1h(r.g1_mul(G1, k))} for k in scalars],
       "g2_mul": [{"k": h32(k), "out": g2h(r.g2_mul(G2, k))} for k in scalars[:8]]}
two_g1 = "030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3" "15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4"
assert kat["g1_mul"][1]["out"] == two_g1, "public 2*G1 vector mismatch"
kat["public_known_answer"] = {"what": "2*G1 on alt_bn128 (EIP-196 ecMul test vector)", "out": two_g1}
pairs = []
for _ in range(2):
    a, b = rng.randrange(r.R), rng.randrange(r.R)
    p, q = r.g1_mul(G1, a), r.g2_mul(G2, b)
    e = r.pairing_lineage(p, q)
    pairs.append({"a": h32(a), "b": h32(b), "p": g1h(p), "q": g2h(q), "e": gth(e)})
kat["pairing_lineage"] = pairs
kx = rng.randrange(r.R)
kat["gt_pow"] = {"base": pairs[0]["e"], "k": h32(kx), "out": gth(r.gt_pow_tower([int(pairs[0]["e"][64 * i:64 * i + 64], 16) for i in range(12)], kx))}
kat["sha3_fr"] = {"A00": h32(r.sha3_fr("A00"))}
assert r.sha3_fr("A00") == 10390014792917408443610864756208359696845607054198933325498802947006247339737
json.dump(kat, open(os.path.join(ROOT, "tests", "golden", "bn254_kat.json"), "w"), indent=1)

# ---- AC17 config 1 through the C++ oracle (reference op sequence)
rng = random.Random(1)
fr = lambda: int(rng.randrange(r.R)).to_bytes(32, "big")
setup_rnd = b"".join(fr() for _ in range(9))
pk, msk = oracle.ac17_setup(setup_rnd)
text, attrs = '("A" and "B") and ("C" and "D")', ["A", "B", "C", "D"]
tree = opol.parse(text, opol.HUMAN)
m, pi, n2 = opol.calculate_msp(tree)
s = fr() + fr()
rho = fr()
msg = oracle.gt_pow(oracle.pairing(oracle.g1_generator(), oracle.g2_generator()), rho)
c0, c, cp = oracle.ac17_cp_encrypt(pk, m, pi, s, msg)
kg_rnd = b"".join(fr() for _ in range(len(attrs) + 3))
k0, k, kp = oracle.ac17_cp_keygen(msk, attrs, kg_rnd)
ok, pruned = opol.calc_pruned(attrs, tree)
dec = oracle.ac17_cp_decrypt([a for a, _ in pruned], pi, c0, c, cp, attrs, k0, k, kp)
assert dec == msg
d = lambda b: hashlib.sha256(bytes(b)).hexdigest()
cfg = {"source": "oracle/ac17.cpp (C++ restatement of ac17/mod.rs:141-430), seeds below; regression + CUDA parity pin",
       "policy": text, "attrs": attrs, "msp": {"m": m, "pi": pi, "n2": n2},
       "setup_rnd": setup_rnd.hex(), "s": s.hex(), "rho": rho.hex(), "keygen_rnd": kg_rnd.hex(),
       "sha256": {"pk": d(pk), "msk": d(msk), "c_0": d(c0), "c": d(c), "c_p": d(cp), "k_0": d(k0), "k": d(k), "k_p": d(kp)},
       "msg": msg.hex()}
json.dump(cfg, open(os.path.join(ROOT, "tests", "golden", "ac17_config1.json"), "w"), indent=1)
print("wrote tests/golden/bn254_kat.json, tests/golden/ac17_config1.json")
