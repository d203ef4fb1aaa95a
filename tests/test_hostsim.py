"""Algorithmic layer of the device headers (tower, curve, Miller loop, final exponentiation)
compiled for the host (tests/hostsim, -DRB_HOST_SIM) and compared with the oracle.  The PTX
Montgomery product is NOT exercised here (portable product instead); tests marked gpu do that."""
import ctypes
import os
import random
import subprocess

import pytest

import oracle
from oracle import pyref as r
from rb_testutil import fr

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    so = os.path.join(HERE, "hostsim", "libhostsim.so")
    deps = [src] + [os.path.join(HERE, "..", "rabe_b200", "csrc", f) for f in ("fp.cuh", "tower.cuh", "tower_body.inc", "curve.cuh", "pairing.cuh", "pairing_body.inc", "consts_gen.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    return ctypes.CDLL(so)


def call(fn, *args, n):
    out = (ctypes.c_uint8 * n)()
    fn(*[bytes(a) for a in args], out)
    return bytes(out)


def test_fields(hs):
    rng = random.Random(5)
    for _ in range(100):
        a, b = rng.randrange(r.P), rng.randrange(r.P)
        assert call(hs.hs_fp_mul, a.to_bytes(32, "big"), b.to_bytes(32, "big"), n=32) == (a * b % r.P).to_bytes(32, "big")
        a, b = rng.randrange(r.R), rng.randrange(r.R)
        assert call(hs.hs_fr_mul, fr(a), fr(b), n=32) == fr(a * b)
    a = rng.randrange(1, r.P)
    assert call(hs.hs_fp_inv, a.to_bytes(32, "big"), n=32) == pow(a, -1, r.P).to_bytes(32, "big")


def test_curves_and_pairing(hs):
    rng = random.Random(6)
    g1, g2 = oracle.g1_generator(), oracle.g2_generator()
    for k in (0, 1, r.R - 1, rng.randrange(r.R), rng.randrange(r.R)):
        assert call(hs.hs_g1_mul, g1, fr(k), n=64) == oracle.g1_mul(g1, fr(k))
        assert call(hs.hs_g2_mul, g2, fr(k), n=128) == oracle.g2_mul(g2, fr(k))
    p, q = oracle.g1_mul(g1, fr(77)), oracle.g1_mul(g1, fr(1234567))
    assert call(hs.hs_g1_add, p, q, n=64) == oracle.g1_add(p, q)
    assert call(hs.hs_g1_add, p, p, n=64) == oracle.g1_add(p, p)
    assert call(hs.hs_g1_add, p, oracle.g1_neg(p), n=64) == b"\0" * 64
    assert hs.hs_on_curve(p, g2) == 3
    q2 = oracle.g2_mul(g2, fr(rng.randrange(r.R)))
    e = call(hs.hs_pairing, p, q2, n=384)
    assert e == oracle.pairing(p, q2)
    assert call(hs.hs_pairing_fixed, p, q2, n=384) == e          # fixed-argument (precomputed lines) path
    # one variable + one fixed pair sharing the Miller accumulator (AC17 decrypt kernel)
    p2, q3 = oracle.g1_mul(g1, fr(rng.randrange(r.R))), oracle.g2_mul(g2, fr(rng.randrange(r.R)))
    assert call(hs.hs_pairing_pair, p, q2, p2, q3, n=384) == oracle.gt_mul(e, oracle.pairing(p2, q3))
    assert call(hs.hs_pairing_pair, p2, q3, p, q2, n=384) == oracle.gt_mul(e, oracle.pairing(p2, q3))
    # the same with the fixed argument's line table normalised to l0 = 1 (loaded AC17 keys): identical Gt value
    assert call(hs.hs_pairing_pair_unit, p, q2, p2, q3, n=384) == oracle.gt_mul(e, oracle.pairing(p2, q3))
    assert call(hs.hs_pairing_pair_unit, p2, q3, p, q2, n=384) == oracle.gt_mul(e, oracle.pairing(p2, q3))
    # four fixed-argument pairs on one accumulator (per-leaf decrypt loops), with and without missing pairs
    ps = [oracle.g1_mul(g1, fr(rng.randrange(r.R))) for _ in range(4)]
    qs = [oracle.g2_mul(g2, fr(rng.randrange(r.R))) for _ in range(4)]
    es = [oracle.pairing(a, b) for a, b in zip(ps, qs)]
    for mask, unit in ((0b1111, 0), (0b0111, 0), (0b0001, 0), (0b1111, 1), (0b0111, 1), (0b0101, 1), (0b0001, 1)):
        exp = oracle.GT_ONE                                  # unit = 1: line tables normalised to l0 = 1 (LSW decrypt)
        for kk in range(4):
            if (mask >> kk) & 1:
                exp = oracle.gt_mul(exp, es[kk])
        out = (ctypes.c_uint8 * 384)()
        hs.hs_pairing_fixed4(b"".join(ps), b"".join(qs), mask, unit, out)
        assert bytes(out) == exp, (mask, unit)
    # the three decrypt terms of an AC17 item on one accumulator
    pv = [oracle.g1_mul(g1, fr(rng.randrange(r.R))) for _ in range(3)]; qv = [oracle.g2_mul(g2, fr(rng.randrange(r.R))) for _ in range(3)]
    pf = [oracle.g1_mul(g1, fr(rng.randrange(r.R))) for _ in range(3)]; qf = [oracle.g2_mul(g2, fr(rng.randrange(r.R))) for _ in range(3)]
    exp = oracle.GT_ONE
    for a, b in list(zip(pv, qv)) + list(zip(pf, qf)):
        exp = oracle.gt_mul(exp, oracle.pairing(a, b))
    out = (ctypes.c_uint8 * 384)()
    hs.hs_pairing_pair3(b"".join(pv), b"".join(qv), b"".join(pf), b"".join(qf), out)
    assert bytes(out) == exp
    k = fr(rng.randrange(r.R))
    assert call(hs.hs_gt_pow, e, k, n=384) == oracle.gt_pow(e, k)
