#!/usr/bin/env python3
"""Writes tests/golden/g2_subgroup.json: points ON the BN254 twist y^2 = x^3 + 3/(9+i), some outside the
order-r subgroup (random x, square root in Fq2), some inside (the same points times the twist cofactor
2p - r, and generator multiples).  `member` is decided by oracle/pyref.py alone with the reference's
decode-time test [r]Q == O (zcash-bn lineage `AffineG2::new`), by plain double-and-add without reducing
the scalar.  tests/ check the C++ oracle (the same [r]Q test) and the CUDA path (the endomorphism test
[u+1]Q + psi([u]Q) + psi^2([u]Q) == psi^3([2u]Q)) against the file."""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref as r

P, R = r.P, r.R


def f2_pow(a, e):
    out = (1, 0)
    while e:
        if e & 1:
            out = r.f2_mul(out, a)
        a = r.f2_mul(a, a); e >>= 1
    return out


def f2_sqrt(a):            # p = 3 mod 4
    a1 = f2_pow(a, (P - 3) // 4)
    alpha = r.f2_mul(r.f2_mul(a1, a1), a)
    x0 = r.f2_mul(a1, a)
    if alpha == (P - 1, 0):
        x = r.f2_mul((0, 1), x0)
    else:
        x = r.f2_mul(f2_pow(r.f2_add((1, 0), alpha), (P - 1) // 2), x0)
    return x if r.f2_mul(x, x) == a else None


def mul_raw(a, k):         # no reduction of k mod r
    acc = None
    while k:
        if k & 1:
            acc = r.g2_add(acc, a)
        a = r.g2_add(a, a); k >>= 1
    return acc


B2 = r.f2_mul((3, 0), r.f2_inv(r.XI))
h32 = lambda x: "%064x" % x
g2h = lambda q: "00" * 128 if q is None else "".join(h32(x) for x in (q[0][0], q[0][1], q[1][0], q[1][1]))
G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781, 11559732032986387107991004021392285783925812861821192530917403151452391805634),
      (8495653923123431417604973247489272438418190587263600148770280649306958101930, 4082367875863433681332203403145435568316851327593401208105741076214120093531))
rng = random.Random(20261018)
pts = [("generator", G2), ("infinity", None), ("generator multiple", r.g2_mul(G2, rng.randrange(R)))]
n = 0
while n < 4:
    x = (rng.randrange(P), rng.randrange(P))
    y = f2_sqrt(r.f2_add(r.f2_mul(r.f2_mul(x, x), x), B2))
    if y is None:
        continue
    q = (x, y)
    assert r.g2_on_curve(q)
    pts.append(("random twist point", q))
    if n < 2:
        pts.append(("the same point times the cofactor 2p - r", mul_raw(q, 2 * P - R)))
    n += 1
out = {"source": "tools/gen_g2_subgroup.py over oracle/pyref.py: member = ([r]Q == O) by plain double-and-add",
       "points": [{"what": w, "q": g2h(q), "member": q is None or mul_raw(q, R) is None} for w, q in pts]}
assert [p["member"] for p in out["points"]] == [True, True, True, False, True, False, True, False, False]
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "g2_subgroup.json"), "w"), indent=1)
print("wrote tests/golden/g2_subgroup.json")
