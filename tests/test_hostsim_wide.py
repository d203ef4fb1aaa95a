"""The six-lane Fq12 layer (rabe_b200/csrc/wide.cuh) compiled for the host -- one host thread per lane, shuffles
as barrier-protected exchanges (tests/hostsim/wide_sim.cpp) -- against the one-thread tower / pairing code and the
oracle.  The PTX carry chains are NOT exercised here (portable twins); tests/test_gpu_wide.py does that."""
import ctypes
import os
import random
import subprocess

import pytest

import oracle
from oracle import pyref as r
from rb_testutil import fr

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "rabe_b200", "csrc")


@pytest.fixture(scope="module")
def ws():
    src = os.path.join(HERE, "hostsim", "wide_sim.cpp")
    so = os.path.join(HERE, "hostsim", "libwidesim.so")
    deps = [src] + [os.path.join(CSRC, f) for f in ("wide.cuh", "fp.cuh", "tower.cuh", "tower_body.inc", "curve.cuh", "pairing.cuh", "pairing_body.inc", "consts_gen.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", so, src])
    return ctypes.CDLL(so)


def gt_rand(rng):
    return b"".join(rng.randrange(r.P).to_bytes(32, "big") for _ in range(12))


def op(ws, fn, code, arg, a, b=None):
    out = (ctypes.c_uint8 * 384)()
    fn(code, arg, bytes(a), bytes(b) if b is not None else None, out)
    return bytes(out)


def test_wide_dot_against_python_ints(ws):
    rng = random.Random(11)
    Rinv = pow(1 << 256, -1, r.P)
    for n in (1, 2, 3, 6):
        for trial in range(20):
            xs = [rng.randrange(r.P) if trial else r.P - 1 for _ in range(n)]
            ys = [rng.randrange(r.P) if trial else r.P - 1 for _ in range(n)]
            out = (ctypes.c_uint8 * 32)()
            ws.ws_dot(b"".join(x.to_bytes(32, "big") for x in xs), b"".join(y.to_bytes(32, "big") for y in ys), n, out)
            # the harness doubles both Montgomery forms WITHOUT reducing (factors up to 2N - 2): recompute exactly
            want = 0
            for x, y in zip(xs, ys):
                xm, ym = x * (1 << 256) % r.P, y * (1 << 256) % r.P
                want += (2 * xm) * (2 * ym)
            want = want * Rinv % r.P * Rinv % r.P            # one Montgomery reduction, then out of Montgomery form
            assert int.from_bytes(bytes(out), "big") == want, (n, trial)


def test_fp12_ops_match_the_one_thread_tower(ws):
    rng = random.Random(12)
    for trial in range(3):
        a, b = gt_rand(rng), gt_rand(rng)
        for code, arg, second in ((0, 0, b), (1, 0, None), (3, 0, None), (4, 1, None), (4, 2, None), (4, 3, None), (5, 0, None), (7, 0, b)):
            assert op(ws, ws.ws_fp12_op, code, arg, a, second) == op(ws, ws.ws_fp12_ref, code, arg, a, second), (code, arg, trial)
    # cyclotomic squaring is only defined on the cyclotomic subgroup: use pairing values
    e = oracle.pairing(oracle.g1_generator(), oracle.g2_generator())
    x = oracle.gt_pow(e, fr(rng.randrange(r.R)))
    assert op(ws, ws.ws_fp12_op, 2, 0, x) == op(ws, ws.ws_fp12_ref, 2, 0, x) == oracle.gt_cyclotomic_sqr(x)
    assert op(ws, ws.ws_fp12_op, 0, 0, x, e) == oracle.gt_mul(x, e)
    assert op(ws, ws.ws_fp12_op, 3, 0, x) == oracle.gt_inverse(x)


def test_final_exponentiation_matches_oracle(ws):
    rng = random.Random(13)
    a = gt_rand(rng)
    got = op(ws, ws.ws_fp12_op, 6, 0, a)
    assert got == op(ws, ws.ws_fp12_ref, 6, 0, a) == oracle.final_exp(a)


def test_three_terms_on_one_accumulator_equal_the_product_of_pairings(ws):
    rng = random.Random(14)
    g1, g2 = oracle.g1_generator(), oracle.g2_generator()
    pv = [oracle.g1_mul(g1, fr(rng.randrange(r.R))) for _ in range(3)]
    pf = [oracle.g1_mul(g1, fr(rng.randrange(r.R))) for _ in range(3)]
    q = [oracle.g2_mul(g2, fr(rng.randrange(r.R))) for _ in range(3)]
    qf = [oracle.g2_mul(g2, fr(rng.randrange(r.R))) for _ in range(3)]
    for mask, unit in ((0b111111, 0), (0b011101, 0), (0b000010, 0), (0, 0), (0b111111, 1), (0b011101, 1), (0b010101, 1)):
        out = (ctypes.c_uint8 * 384)()                    # unit = 1: fixed-argument line tables normalised to l0 = 1 (loaded keys)
        ws.ws_pairing_terms(b"".join(pv), b"".join(q), b"".join(pf), b"".join(qf), mask, unit, out)
        want = oracle.GT_ONE
        for j in range(3):
            if (mask >> (2 * j)) & 1:
                want = oracle.gt_mul(want, oracle.pairing(pv[j], q[j]))
            if (mask >> (2 * j + 1)) & 1:
                want = oracle.gt_mul(want, oracle.pairing(pf[j], qf[j]))
        assert bytes(out) == want, (bin(mask), unit)
