"""The six-lane pairing layer (rabe_b200/csrc/wide.cuh) on the device: the PTX carry chains of the wide
accumulator against Python integers, every Fq12 operation and the final exponentiation against the oracle, the
AC17 decrypt kernel (three terms of a ciphertext on one Miller accumulator) against the two-lane kernels and the
oracle, including points at infinity and batch sizes that do not fill a warp (five work items per warp)."""
import ctypes
import random

import numpy as np
import pytest

import oracle
from oracle import policy as opol
from oracle import pyref as r
import rb_testutil as util
from rb_testutil import fr, rand_fr, u8

pytestmark = pytest.mark.gpu


def _call(engine, name, *args):
    from rabe_b200._lib import check
    check(getattr(engine.L, name)(engine.ctx, *args), name)


def gt_rand(rng):
    return b"".join(rng.randrange(r.P).to_bytes(32, "big") for _ in range(12))


def w6_op(engine, op, arg, a, b):
    n = len(a) // 384
    A, B, out = u8(a), u8(b), np.empty(n * 384, dtype=np.uint8)
    _call(engine, "rb_dbg_w6_op", op, arg, ctypes.c_void_p(A.ctypes.data), ctypes.c_void_p(B.ctypes.data), n, ctypes.c_void_p(out.ctypes.data))
    return out.tobytes()


def test_wide_accumulator_carry_chains(engine):
    rng = random.Random(21)
    Rinv = pow(1 << 256, -1, r.P)
    for K in (1, 2, 3, 6):
        n = 257
        xs = [[rng.randrange(r.P) for _ in range(K)] for _ in range(n)]
        ys = [[rng.randrange(r.P) for _ in range(K)] for _ in range(n)]
        xs[0] = [r.P - 1] * K; ys[0] = [r.P - 1] * K                      # the largest operands
        xs[1] = [0] * K; ys[2] = [0] * K
        xs[3] = [(1 << 253) + 12345] * K; ys[3] = [r.P - 2] * K
        X = u8(b"".join(v.to_bytes(32, "big") for row in xs for v in row)); Y = u8(b"".join(v.to_bytes(32, "big") for row in ys for v in row))
        out = np.empty(n * 32, dtype=np.uint8)
        _call(engine, "rb_dbg_wide_dot", ctypes.c_void_p(X.ctypes.data), ctypes.c_void_p(Y.ctypes.data), K, n, ctypes.c_void_p(out.ctypes.data))
        got = out.tobytes()
        for i in range(n):
            want = sum((2 * (x * (1 << 256) % r.P)) * (2 * (y * (1 << 256) % r.P)) for x, y in zip(xs[i], ys[i])) * Rinv % r.P * Rinv % r.P
            assert int.from_bytes(got[32 * i:32 * i + 32], "big") == want, (K, i)


def test_dedicated_squaring_carry_chains(engine):
    """fp.cuh fe_sqr (28 cross products + 8 squares + one 512-bit Montgomery reduction) against Python integers: Fq, Fr, and
    Fq with an unreduced operand (a + b < 2N), on random values and the edge patterns that stress every carry target."""
    rng = random.Random(22)
    for mode, mod in ((0, r.P), (1, r.R), (2, r.P)):
        edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 253) - 1, 1 << 253, (1 << 224) - 1, 0xffffffff, 0xffffffff << 32,
                int("ffffffff00000000" * 3 + "0fffffff00000000", 16) % mod, int("00000000ffffffff" * 4, 16) % mod]
        xs = edge + [rng.randrange(mod) for _ in range(1024 - len(edge))]
        ys = [mod - 1] * len(edge) + [rng.randrange(mod) for _ in range(1024 - len(edge))]
        n = len(xs)
        A = u8(b"".join(v.to_bytes(32, "big") for v in xs)); B = u8(b"".join(v.to_bytes(32, "big") for v in ys))
        out = np.empty(n * 32, dtype=np.uint8)
        _call(engine, "rb_dbg_fq_sqr", ctypes.c_void_p(A.ctypes.data), ctypes.c_void_p(B.ctypes.data), mode, n, ctypes.c_void_p(out.ctypes.data))
        got = out.tobytes()
        for i in range(n):
            v = xs[i] + ys[i] if mode == 2 else xs[i]
            assert int.from_bytes(got[32 * i:32 * i + 32], "big") == v * v % mod, (mode, i)


def test_fp12_operations_against_oracle(engine):
    rng = random.Random(22)
    for n in (1, 4, 5, 6, 23):                                            # partial warps, exact warps, several warps
        a = [gt_rand(rng) for _ in range(n)]; b = [gt_rand(rng) for _ in range(n)]
        A, Bb = b"".join(a), b"".join(b)
        assert w6_op(engine, 0, 0, A, Bb) == b"".join(oracle.gt_mul(x, y) for x, y in zip(a, b))
        assert w6_op(engine, 1, 0, A, A) == b"".join(oracle.gt_mul(x, x) for x in a)
        assert w6_op(engine, 3, 0, A, A) == b"".join(oracle.gt_inverse(x) for x in a)
        for j in (1, 2, 3):
            assert w6_op(engine, 4, j, A, A) == b"".join(oracle.gt_frobenius(x, j) for x in a)
        assert w6_op(engine, 6, 0, A, A) == b"".join(oracle.final_exp(x) for x in a)
    e = oracle.pairing(oracle.g1_generator(), oracle.g2_generator())
    cyc = [oracle.gt_pow(e, fr(rng.randrange(r.R))) for _ in range(7)]
    C = b"".join(cyc)
    assert w6_op(engine, 2, 0, C, C) == b"".join(oracle.gt_cyclotomic_sqr(x) for x in cyc)
    # sparse line product against the dense product with the same element: l0 + l3 w^3 + l4 w^4 in the tower layout
    l = [[rng.randrange(r.P) for _ in range(6)] for _ in range(7)]
    lines = b"".join(b"".join(v.to_bytes(32, "big") for v in row) + b"\0" * 192 for row in l)
    dense = []
    for row in l:
        c = [0] * 12
        c[0], c[1] = row[0], row[1]                # w^0 -> tower index 0
        c[8], c[9] = row[2], row[3]                # w^3 -> tower index 4 (c1.c1)
        c[4], c[5] = row[4], row[5]                # w^4 -> tower index 2 (c0.c2)
        dense.append(b"".join(v.to_bytes(32, "big") for v in c))
    a = [gt_rand(rng) for _ in range(7)]
    assert w6_op(engine, 7, 0, b"".join(a), lines) == b"".join(oracle.gt_mul(x, d) for x, d in zip(a, dense))


def test_ac17_decrypt_six_lane_equals_two_lane_and_oracle(engine):
    """B = 1, 4, 5, 6, 11 ciphertexts (five items per warp), a loaded key and an inline key, and items whose
    sums hit the point at infinity -- against the two-lane kernels and, element-wise, the oracle."""
    rng = random.Random(23)
    pk, msk = oracle.ac17_setup(rand_fr(rng, 9))
    names = ["A", "B", "C", "D"]
    policy = '("A" and "B") and ("C" or "D")'
    tree = opol.parse(policy, opol.HUMAN)
    m, pi, n2 = opol.calculate_msp(tree)
    k0, k, kp = oracle.ac17_cp_keygen(msk, names, rand_fr(rng, len(names) + 3))
    ok, pruned = opol.calc_pruned(names, tree)
    ct_idx, sk_idx = util.decrypt_lists(pruned, pi, names)
    plist = [a for a, _ in pruned]
    for B in (1, 4, 5, 6, 11):
        cts = [oracle.ac17_cp_encrypt(pk, m, pi, rand_fr(rng, 2), util.gt_random(rng)) for _ in range(B)]
        c0, c, cp = (b"".join(x[i] for x in cts) for i in range(3))
        want = b"".join(oracle.ac17_cp_decrypt(plist, pi, x[0], x[1], x[2], names, k0, k, kp) for x in cts)
        got = {}
        for six in (2, 1):
            _call(engine, "rb_ctx_set_pairing_layout", six)
            try:
                got[six] = engine.ac17_cp_decrypt(u8(k0), u8(k), u8(kp), u8(c0), u8(c), u8(cp), len(pi), ct_idx, sk_idx).tobytes()
                skh = engine.ac17_sk_load(u8(k0), u8(k), u8(kp))
                assert engine.ac17_cp_decrypt_sk(skh, u8(c0), u8(c), u8(cp), len(pi), ct_idx, sk_idx).tobytes() == got[six]
            finally:
                _call(engine, "rb_ctx_set_pairing_layout", 0)
        assert got[2] == got[1] == want, B
    # points at infinity: c_0 members at infinity, an empty gather list (prod_g = infinity), k_p = infinity
    x = cts[0]
    c0_inf = x[0][:128] + b"\0" * 128 + x[0][256:]
    for six in (2, 1):
        _call(engine, "rb_ctx_set_pairing_layout", six)
        try:
            a1 = engine.ac17_cp_decrypt(u8(k0), u8(k), u8(kp), u8(c0_inf), u8(x[1]), u8(x[2]), len(pi), ct_idx, sk_idx).tobytes()
            a2 = engine.ac17_cp_decrypt(u8(k0), u8(k), u8(b"\0" * 192), u8(x[0]), u8(x[1]), u8(x[2]), len(pi), [], []).tobytes()
        finally:
            _call(engine, "rb_ctx_set_pairing_layout", 0)
        if six == 2:
            keep = (a1, a2)
        else:
            assert (a1, a2) == keep
    assert keep[1] == x[2]                       # no pairs at all: msg_out = c_p


def test_pairing_product_final_exponentiation_six_lane(engine):
    """rb_pairing_product_batch: Miller values from the two-lane kernel, product + final exponentiation on six lanes;
    ragged lists (0, 1, 3, 7 pairs) in one call."""
    rng = random.Random(24)
    g1, g2 = oracle.g1_generator(), oracle.g2_generator()
    sizes = [0, 1, 3, 7, 2, 1, 0, 5]
    P, Q, offs, want = b"", b"", [0], []
    for sz in sizes:
        acc = oracle.GT_ONE
        for _ in range(sz):
            p, q = oracle.g1_mul(g1, fr(rng.randrange(r.R))), oracle.g2_mul(g2, fr(rng.randrange(r.R)))
            P += p; Q += q
            acc = oracle.gt_mul(acc, oracle.pairing(p, q))
        offs.append(offs[-1] + sz); want.append(acc)
    assert engine.pairing_product(u8(P), u8(Q), offs).tobytes() == b"".join(want)
