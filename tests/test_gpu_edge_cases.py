"""Edge cases of the C ABI on the GPU: empty batches, sizes that are not multiples of the
per-thread run (16) or the block size, invalid inputs (status codes, never aborts), points at
infinity, per-item gather lists, device-resident buffers, and reuse of a context after errors."""
import random

import numpy as np
import pytest

import oracle
from oracle import policy as opol
import rb_testutil as util
from rb_testutil import R, fr, rand_fr, u8

pytestmark = pytest.mark.gpu


def test_empty_batches_are_ok(engine):
    g = oracle.g1_generator()
    tab = engine.g1_table(u8(g), 8)
    e = np.empty(0, dtype=np.uint8)
    assert engine.g1_mul_fixed(tab, e).size == 0
    assert engine.fq_mul(e, e).size == 0
    assert engine.g1_mul_var(e, e).size == 0
    assert engine.pairing_product(e, e, [0]).size == 0
    one = np.zeros(1, dtype=np.uint8)                                          # non-null placeholder buffers
    assert engine.pairing_product(one[:0], one[:0], [0, 0]).tobytes() == oracle.GT_ONE   # one empty product = 1


def test_ragged_sizes_g1_fixed(engine):
    """n not a multiple of 16 (outputs per thread) nor of 128*16 (per block)."""
    rng = random.Random(51)
    g = oracle.g1_mul(oracle.g1_generator(), fr(rng.randrange(R)))
    tab = engine.g1_table(u8(g), 16)
    for n in (1, 15, 16, 17, 2049):
        ks = [rng.randrange(R) for _ in range(n)]
        out = engine.g1_mul_fixed(tab, u8(b"".join(fr(k) for k in ks))).tobytes()
        for i in sorted(set([0, n // 2, n - 1])):
            assert out[64 * i:64 * i + 64] == oracle.g1_mul(g, fr(ks[i])), (n, i)


def test_zero_scalars_give_infinity_inside_a_batch(engine):
    """An infinite result inside a Montgomery-trick run must not poison its neighbours."""
    g = oracle.g1_generator()
    tab = engine.g1_table(u8(g), 16)
    ks = [5, 0, 7, 0, 0, 11] + [0] * 10 + [13]
    out = engine.g1_mul_fixed(tab, u8(b"".join(fr(k) for k in ks))).tobytes()
    for i, k in enumerate(ks):
        assert out[64 * i:64 * i + 64] == (b"\0" * 64 if k == 0 else oracle.g1_mul(g, fr(k))), i


def test_invalid_inputs_return_status_codes(engine):
    from rabe_b200._lib import RB_EINVAL, RB_ENOTMEMBER, RB_EPOLICY, RabeB200Error
    g = oracle.g1_generator()
    with pytest.raises(RabeB200Error) as ei:                       # scalar >= r
        engine.g1_mul_var(u8(g), u8(R.to_bytes(32, "big")))
    assert ei.value.status == RB_ENOTMEMBER
    with pytest.raises(RabeB200Error) as ei:                       # coordinate >= p
        engine.g1_mul_var(u8(b"\xff" * 64), u8(fr(1)))
    assert ei.value.status == RB_ENOTMEMBER
    bad_g2 = bytearray(oracle.g2_generator()); bad_g2[127] ^= 1
    with pytest.raises(RabeB200Error) as ei:                       # G2 point off the twist
        engine.pairing(u8(g), u8(bytes(bad_g2)))
    assert ei.value.status == RB_ENOTMEMBER
    with pytest.raises(RabeB200Error) as ei:                       # unsupported window
        engine.g1_table(u8(g), 3)
    assert ei.value.status == RB_EINVAL
    with pytest.raises(RabeB200Error) as ei:                       # MSP entry outside {-1,0,1}
        engine.msp_load(np.array([[2]], dtype=np.int8), u8(b"\0" * 192), u8(b"\0" * 192))
    assert ei.value.status == RB_EPOLICY
    with pytest.raises(RabeB200Error) as ei:                       # inverse of zero
        engine.fr_op("inverse", u8(fr(0)))
    assert ei.value.status == RB_ENOTMEMBER
    # the context is still healthy
    assert engine.g1_mul_var(u8(g), u8(fr(3))).tobytes() == oracle.g1_mul(g, fr(3))


def test_decrypt_index_out_of_range_rejected(engine):
    from rabe_b200._lib import RB_EINVAL, RabeB200Error
    z = lambda n: np.zeros(n, dtype=np.uint8)
    with pytest.raises(RabeB200Error) as ei:
        engine.ac17_cp_decrypt(z(384), z(192), z(192), z(384), z(192), z(384), 1, [1], [0])
    assert ei.value.status == RB_EINVAL


def test_pairing_with_infinity_and_repeated_points(engine):
    g, h = oracle.g1_generator(), oracle.g2_generator()
    inf1, inf2 = b"\0" * 64, b"\0" * 128
    e = oracle.pairing(g, h)
    out = engine.pairing_product(u8(g + inf1 + g), u8(h + h + inf2), [0, 3]).tobytes()
    assert out == e
    out = engine.pairing_product(u8(g + g), u8(h + h), [0, 2]).tobytes()
    assert out == oracle.gt_mul(e, e)
    out = engine.pairing_product(u8(g + oracle.g1_neg(g)), u8(h + h), [0, 2]).tobytes()
    assert out == oracle.GT_ONE


def test_gather_sum_special_cases(engine):
    """Doubling, cancellation to infinity and infinity inputs inside sums (complete addition)."""
    g = oracle.g1_generator()
    p = oracle.g1_mul(g, fr(9)); n = oracle.g1_neg(p)
    pts = p + n + b"\0" * 64
    lists = [[0, 0, 0], [0, 1], [2, 0, 2], [0, 1, 0], [2, 2]]
    idx = [i for l in lists for i in l]; offs = np.cumsum([0] + [len(l) for l in lists])
    out = engine.g1_sum_gather(u8(pts), idx, offs).tobytes()
    exp = [oracle.g1_mul(g, fr(27)), b"\0" * 64, p, p, b"\0" * 64]
    for i, e in enumerate(exp):
        assert out[64 * i:64 * i + 64] == e, i


def test_ac17_per_item_lists_and_device_buffers(engine):
    """ct_offs / sk_offs (one gather list per item) and torch CUDA tensors as buffers."""
    import torch
    rng = random.Random(53)
    pk, msk = oracle.ac17_setup(rand_fr(rng, 9))
    policy = '("A" or "B") and ("C" or "D")'
    tree = opol.parse(policy, opol.HUMAN)
    m, pi, n2 = opol.calculate_msp(tree)
    h_row, h_col = util.ac17_hashes(pi, n2)
    pkh = engine.ac17_pk_load(u8(pk)); msp = engine.msp_load(np.array(m, dtype=np.int8), u8(h_row), u8(h_col))
    attrs = ["A", "B", "C", "D"]
    k0, k, kp = oracle.ac17_cp_keygen(msk, attrs, rand_fr(rng, len(attrs) + 3))
    B = 4
    msgs = [util.gt_random(rng) for _ in range(B)]
    s = rand_fr(rng, 2 * B)
    dev = torch.device("cuda:0")
    t = lambda b: torch.from_numpy(u8(b)).to(dev)
    c0, c, cp = engine.ac17_cp_encrypt(pkh, msp, t(s), t(b"".join(msgs)))
    engine.status()
    assert c0.is_cuda and c.is_cuda
    # item b uses a different satisfying set: {A,C}, {A,D}, {B,C}, {B,D}
    sets = [["A", "C"], ["A", "D"], ["B", "C"], ["B", "D"]]
    ct_idx, sk_idx, offs = [], [], [0]
    for sset in sets:
        ct_idx += [pi.index(a) for a in sset]; sk_idx += [attrs.index(a) for a in sset]; offs.append(len(ct_idx))
    out = engine.ac17_cp_decrypt(t(k0), t(k), t(kp), c0, c, cp, len(pi), np.array(ct_idx, dtype=np.uint32), np.array(sk_idx, dtype=np.uint32),
                                 ct_offs=np.array(offs, dtype=np.uint32), sk_offs=np.array(offs, dtype=np.uint32))
    engine.status()
    assert out.cpu().numpy().tobytes() == b"".join(msgs)
    skh = engine.ac17_sk_load(u8(k0), u8(k), u8(kp))
    out2 = engine.ac17_cp_decrypt_sk(skh, c0, c, cp, len(pi), np.array(ct_idx, dtype=np.uint32), np.array(sk_idx, dtype=np.uint32),
                                     ct_offs=np.array(offs, dtype=np.uint32), sk_offs=np.array(offs, dtype=np.uint32))
    engine.status()
    assert out2.cpu().numpy().tobytes() == b"".join(msgs)
    # a wrong list decrypts to something else, bit-exactly what the reference loops would produce
    ref = oracle.ac17_cp_decrypt(["A"], pi, c0.cpu().numpy().tobytes()[:384], c.cpu().numpy().tobytes()[:192 * len(pi)],
                                 cp.cpu().numpy().tobytes()[:384], attrs, k0, k, kp)
    got = engine.ac17_cp_decrypt(u8(k0), u8(k), u8(kp), c0[:384], c[:192 * len(pi)], cp[:384], len(pi), [pi.index("A")], [attrs.index("A")])
    engine.status()
    assert got.cpu().numpy().tobytes() == ref != msgs[0]


def test_two_contexts_are_independent(engine):
    from rabe_b200.engine import Engine
    other = Engine(0)
    g = oracle.g1_generator()
    a = engine.g1_mul_var(u8(g), u8(fr(123)))
    b = other.g1_mul_var(u8(g), u8(fr(123)))
    assert a.tobytes() == b.tobytes() == oracle.g1_mul(g, fr(123))
    other.close()
    assert engine.g1_mul_var(u8(g), u8(fr(5))).tobytes() == oracle.g1_mul(g, fr(5))


def test_fused_scheme_entry_points_reject_bad_indices_and_accept_empty_batches(engine):
    import ctypes
    from rabe_b200._lib import RB_EINVAL, RB_OK
    L, ctx = engine.L, engine.ctx
    z32, z64, z128, z384 = (np.zeros(n, dtype=np.uint8) for n in (32, 64, 128, 384))
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    bad = np.array([7], dtype=np.uint32)                       # index 7 with n = n_k = 1
    ok0 = np.array([0], dtype=np.uint32)
    assert L.rb_bsw_decrypt_batch(ctx, p(z128), p(z64), p(z128), 1, p(z64), p(z384), p(z64), p(z128), 1, p(bad), p(ok0), p(z32), 1, 1, p(z384)) == RB_EINVAL
    assert L.rb_bsw_decrypt_batch(ctx, p(z128), p(z64), p(z128), 1, p(z64), p(z384), p(z64), p(z128), 1, p(ok0), p(bad), p(z32), 1, 1, p(z384)) == RB_EINVAL
    assert L.rb_lsw_decrypt_batch(ctx, p(z64), p(z128), 1, p(z384), p(z128), p(z64), 1, p(bad), p(ok0), p(z32), 1, 1, p(z384)) == RB_EINVAL
    # B == 0 is a no-op, null pointers are RB_EINVAL
    assert L.rb_bsw_decrypt_batch(ctx, p(z128), p(z64), p(z128), 1, p(z64), p(z384), p(z64), p(z128), 1, p(ok0), p(ok0), p(z32), 1, 0, p(z384)) == RB_OK
    assert L.rb_bsw_decrypt_batch(ctx, None, p(z64), p(z128), 1, p(z64), p(z384), p(z64), p(z128), 1, p(ok0), p(ok0), p(z32), 1, 1, p(z384)) == RB_EINVAL
    assert L.rb_sha3_fr_batch(ctx, None, p(ok0), 0, p(z32)) == RB_OK
    # all-infinity inputs decrypt to c_p * 1 (every pair is masked), not to an error or a hang
    one = bytearray(384); one[31] = 1
    cp = np.frombuffer(bytes(one), dtype=np.uint8).copy()
    out = np.empty(384, dtype=np.uint8)
    st = L.rb_lsw_decrypt_batch(ctx, p(z64), p(z128), 1, p(cp), p(z128), p(z64), 1, p(ok0), p(ok0), p(z32), 1, 1, p(out))
    assert st != RB_OK or out.tobytes() == bytes(one)          # an infinite key point is rejected (line tables) or masked
    engine.status() if st == RB_OK else None


def test_policy_reload_and_packed_sha3_edge_cases(engine):
    """rb_msp_reload_batch / rb_sha3_fr_batch_len: null arguments, a stated data length that contradicts
    the offsets, matrix entries outside {-1,0,1}, and the empty string (SHA3-256("") mod r)."""
    import ctypes
    import hashlib
    from rabe_b200._lib import RB_EINVAL, RB_EPOLICY, RB_OK
    L, ctx = engine.L, engine.ctx
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    data = np.frombuffer(b"A00", dtype=np.uint8).copy()
    offs = np.array([0, 3, 3], dtype=np.uint32)                # "A00" and ""
    out = np.zeros(64, dtype=np.uint8)
    assert L.rb_sha3_fr_batch_len(ctx, p(data), 4, p(offs), 2, p(out)) == RB_EINVAL       # offs[n] = 3, not 4
    assert L.rb_sha3_fr_batch_len(ctx, p(data), 3, None, 2, p(out)) == RB_EINVAL
    assert L.rb_sha3_fr_batch_len(ctx, p(data), 3, p(offs), 2, p(out)) == RB_OK
    assert out[:32].tobytes() == oracle.sha3_fr("A00")
    assert int.from_bytes(out[32:].tobytes(), "big") == int.from_bytes(hashlib.sha3_256(b"").digest(), "big") % R
    m = np.zeros((2, 1, 1), dtype=np.int8)
    h = np.zeros(2 * 6 * 32, dtype=np.uint8)
    msp = engine.msp_load_batch(m, h, h)
    assert L.rb_msp_reload_batch(ctx, msp.ptr, None, p(h), p(h), 0) == RB_EINVAL
    assert L.rb_msp_reload_batch(ctx, None, p(m.view(np.uint8)), p(h), p(h), 0) == RB_EINVAL
    bad = np.full((2, 1, 1), 2, dtype=np.int8)
    assert L.rb_msp_reload_batch(ctx, msp.ptr, p(bad.view(np.uint8)), p(h), p(h), 0) == RB_EPOLICY
    assert L.rb_msp_reload_batch(ctx, msp.ptr, p(m.view(np.uint8)), p(h), p(h), 0) == RB_OK
    assert L.rb_msp_reload_batch(ctx, msp.ptr, p(m.view(np.uint8)), p(h), p(h[:192]), 1) == RB_OK
    engine.status()
