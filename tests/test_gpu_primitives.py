"""GPU parity of the L0 operators (through the C ABI) against the CPU oracle / Python integers."""
import random

import numpy as np
import pytest

import oracle
from rb_testutil import P, R, fr, rand_fr, u8, gt_random

pytestmark = pytest.mark.gpu


def test_fq_fr_mul(engine):
    rng = random.Random(11)
    n = 4096
    edge = [0, 1, 2, P - 1, P - 2, (1 << 253), (1 << 254) - 1 if (1 << 254) - 1 < P else P - 3]
    a = edge + [rng.randrange(P) for _ in range(n - len(edge))]
    b = list(reversed(edge)) + [rng.randrange(P) for _ in range(n - len(edge))]
    out = engine.fq_mul(u8(b"".join(x.to_bytes(32, "big") for x in a)), u8(b"".join(x.to_bytes(32, "big") for x in b))).tobytes()
    for i in range(n):
        assert out[32 * i:32 * i + 32] == (a[i] * b[i] % P).to_bytes(32, "big"), i
    a = [x % R for x in a]; b = [x % R for x in b]
    out = engine.fr_mul(u8(b"".join(fr(x) for x in a)), u8(b"".join(fr(x) for x in b))).tobytes()
    for i in range(n):
        assert out[32 * i:32 * i + 32] == fr(a[i] * b[i]), i


def test_fq_rejects_unreduced(engine):
    from rabe_b200._lib import RabeB200Error
    bad = u8(P.to_bytes(32, "big"))
    with pytest.raises(RabeB200Error):
        engine.fq_mul(bad, bad)


def test_g1_fixed_and_var(engine):
    rng = random.Random(12)
    g = oracle.g1_mul(oracle.g1_generator(), fr(rng.randrange(R)))
    ks = [0, 1, 2, R - 1, R - 2, 65535, 65536, 1 << 240] + [rng.randrange(R) for _ in range(41)]
    kb = b"".join(fr(k) for k in ks)
    for w in (16, 8, 5, 18):                  # 18: chunked large-window table builder
        tab = engine.g1_table(g, w)
        out = engine.g1_mul_fixed(tab, u8(kb)).tobytes()
        for i, k in enumerate(ks):
            assert out[64 * i:64 * i + 64] == oracle.g1_mul(g, fr(k)), (w, i)
        tab.close()
    pts = b"".join(oracle.g1_mul(g, fr(rng.randrange(R))) for _ in ks)
    out = engine.g1_mul_var(u8(pts), u8(kb)).tobytes()
    for i, k in enumerate(ks):
        assert out[64 * i:64 * i + 64] == oracle.g1_mul(pts[64 * i:64 * i + 64], fr(k)), i


def test_g1_fixed_largest_windows(engine):
    """The bench's table widths: 24 bits (11 windows, 11.8 GB) and 26 bits (10 windows, 42.9 GB), bit-exact against
    the oracle on scalars that sit on the window boundaries."""
    from rabe_b200._lib import RabeB200Error
    rng = random.Random(14)
    g = oracle.g1_mul(oracle.g1_generator(), fr(rng.randrange(R)))
    for w in (24, 26):
        last = (253 // w) * w
        ks = [0, 1, R - 1, (1 << w) - 1, 1 << w, (1 << (2 * w)) - 1, 1 << last, (1 << last) - 1, (R - 1) >> 1] + [rng.randrange(R) for _ in range(23)]
        tab = engine.g1_table(g, w)
        out = engine.g1_mul_fixed(tab, u8(b"".join(fr(k) for k in ks))).tobytes()
        tab.close()
        for i, k in enumerate(ks):
            assert out[64 * i:64 * i + 64] == oracle.g1_mul(g, fr(k)), (w, i)
    with pytest.raises(RabeB200Error):
        engine.g1_table(g, 27)


def test_g1_fixed_grid_shapes(engine):
    """The launch picks the outputs per thread (2..24) so that the grid is whole waves of resident threads: every batch
    size -- one output, a ragged tail, a little over one and over two waves -- must give the same bytes as the oracle."""
    import numpy as np
    rng = random.Random(15)
    g = oracle.g1_mul(oracle.g1_generator(), fr(rng.randrange(R)))
    base = [0, 1, R - 1, 1 << 16, (1 << 16) - 1, 1 << 32] + [rng.randrange(R) for _ in range(44)]
    want = np.frombuffer(b"".join(oracle.g1_mul(g, fr(k)) for k in base), dtype=np.uint8).reshape(len(base), 64)
    kb = np.frombuffer(b"".join(fr(k) for k in base), dtype=np.uint8).reshape(len(base), 32)
    tab = engine.g1_table(g, 16)
    for n in (1, 2, 3, 5, 33, 1000, 37889, 2 * 37888 + 5, 24 * 37888 + 1):
        idx = np.arange(n) % len(base)
        out = engine.g1_mul_fixed(tab, np.ascontiguousarray(kb[idx]).reshape(-1)).reshape(n, 64)
        assert (np.asarray(out) == want[idx]).all(), n
    tab.close()


def test_g2_fixed_and_var(engine):
    rng = random.Random(13)
    h = oracle.g2_mul(oracle.g2_generator(), fr(rng.randrange(R)))
    ks = [0, 1, R - 1, 255, 256] + [rng.randrange(R) for _ in range(12)]
    kb = b"".join(fr(k) for k in ks)
    for w in (8, 13):                         # 13: chunked large-window table builder
        tab = engine.g2_table(h, w)
        out = engine.g2_mul_fixed(tab, u8(kb)).tobytes()
        for i, k in enumerate(ks):
            assert out[128 * i:128 * i + 128] == oracle.g2_mul(h, fr(k)), (w, i)
        tab.close()
    out = engine.g2_mul_var(u8(h * len(ks)), u8(kb)).tobytes()
    for i, k in enumerate(ks):
        assert out[128 * i:128 * i + 128] == oracle.g2_mul(h, fr(k)), i


def test_gt_ops(engine):
    rng = random.Random(14)
    a, b = gt_random(rng), gt_random(rng)
    ks = [0, 1, R - 1] + [rng.randrange(R) for _ in range(5)]
    kb = b"".join(fr(k) for k in ks)
    for w in (8, 13):                         # 13: chunked large-window table builder
        tab = engine.gt_table(a, w)
        out = engine.gt_pow_fixed(tab, u8(kb)).tobytes()
        for i, k in enumerate(ks):
            assert out[384 * i:384 * i + 384] == oracle.gt_pow(a, fr(k)), (w, i)
        tab.close()
    out = engine.gt_pow_var(u8(a * len(ks)), u8(kb)).tobytes()
    for i, k in enumerate(ks):
        assert out[384 * i:384 * i + 384] == oracle.gt_pow(a, fr(k)), i
    assert engine.gt_mul(u8(a), u8(b)).tobytes() == oracle.gt_mul(a, b)
    assert engine.gt_inverse(u8(a)).tobytes() == oracle.gt_inverse(a)


def test_gather_sum(engine):
    rng = random.Random(15)
    g = oracle.g1_generator()
    pts = [oracle.g1_mul(g, fr(rng.randrange(R))) for _ in range(10)]
    pts.append(oracle.g1_neg(pts[0]))            # index 10 = -pts[0]
    lists = [[0, 1, 2], [], [3], [0, 0], [0, 10], [4, 5, 6, 7, 8, 9, 4]]
    idx = [i for l in lists for i in l]
    offs = np.cumsum([0] + [len(l) for l in lists])
    out = engine.g1_sum_gather(u8(b"".join(pts)), idx, offs).tobytes()
    for o, l in enumerate(lists):
        exp = b"\0" * 64
        for i in l:
            exp = oracle.g1_add(exp, pts[i])
        assert out[64 * o:64 * o + 64] == exp, o


def test_pairing_products(engine):
    rng = random.Random(16)
    g, h = oracle.g1_generator(), oracle.g2_generator()
    Ps = [oracle.g1_mul(g, fr(rng.randrange(R))) for _ in range(7)] + [b"\0" * 64]
    Qs = [oracle.g2_mul(h, fr(rng.randrange(R))) for _ in range(8)]
    single = engine.pairing(u8(b"".join(Ps)), u8(b"".join(Qs))).tobytes()
    refs = [oracle.pairing(p, q) for p, q in zip(Ps, Qs)]
    for i in range(8):
        assert single[384 * i:384 * i + 384] == refs[i], i
    assert refs[7] == oracle.GT_ONE
    offs = [0, 3, 3, 8]
    prod = engine.pairing_product(u8(b"".join(Ps)), u8(b"".join(Qs)), offs).tobytes()
    exp0 = oracle.gt_mul(oracle.gt_mul(refs[0], refs[1]), refs[2])
    exp2 = oracle.GT_ONE
    for i in range(3, 8):
        exp2 = oracle.gt_mul(exp2, refs[i])
    assert prod[:384] == exp0 and prod[384:768] == oracle.GT_ONE and prod[768:] == exp2


def test_not_on_curve_rejected(engine):
    from rabe_b200._lib import RabeB200Error
    bad = bytearray(oracle.g1_generator()); bad[63] ^= 1
    with pytest.raises(RabeB200Error):
        engine.g1_mul_var(u8(bytes(bad)), u8(fr(5)))
    # the context stays usable afterwards
    assert engine.g1_mul_var(u8(oracle.g1_generator()), u8(fr(5))).tobytes() == oracle.g1_mul(oracle.g1_generator(), fr(5))


def test_sha3_fr_on_device(engine):
    """hash/mod.rs:23-31 on the device: SHA3-256 -> big-endian integer mod r, against hashlib."""
    import hashlib
    rng = random.Random(17)
    msgs = [b"", b"A00", b"a6301", "0641".encode(), b"x" * 135, b"y" * 136, b"z" * 137, b"w" * 272, b"v" * 500] + \
           [bytes(rng.randrange(256) for _ in range(rng.randrange(1, 300))) for _ in range(40)]
    out = engine.sha3_fr(msgs).tobytes()
    for i, m in enumerate(msgs):
        assert out[32 * i:32 * i + 32] == fr(int.from_bytes(hashlib.sha3_256(m).digest(), "big") % R), i
    assert out[32:64] == oracle.sha3_fr("A00")
