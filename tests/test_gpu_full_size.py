"""BASELINE.json configurations 2-5 at their full attribute counts, checked through
size-independent properties (encrypt -> decrypt round trips, wrong-key failure) because the CPU
oracle needs minutes per item at these sizes.  Element-wise parity at small sizes is in
test_gpu_ac17.py / test_gpu_schemes.py."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PLAINTEXT = b"dance like no one's watching, encrypt like everyone is!"


@pytest.fixture(scope="module")
def mods(engine):
    from rabe_b200.schemes import ac17, aw11, bsw, common, lsw
    from rabe_b200.policy import PolicyLanguage
    common.set_engine(engine)
    return ac17, bsw, lsw, aw11, common, PolicyLanguage


def and_tree(groups):
    """root AND over len(groups) subtrees, each an AND over its leaves (n-ary gates)."""
    return "(" + " and ".join("(" + " and ".join('"%s"' % a for a in g) + ")" for g in groups) + ")"


def binary_tree(names, rng, p_and=0.5):
    if len(names) == 1:
        return '"%s"' % names[0]
    k = rng.randrange(1, len(names))
    op = "and" if rng.random() < p_and else "or"
    return "(%s %s %s)" % (binary_tree(names[:k], rng, p_and), op, binary_tree(names[k:], rng, p_and))


def test_config2_ac17_batch_512_round_trip(mods):
    """AC17 CP, 64 attributes, random binary AND/OR policy, a 512-item batch through the API mirror."""
    ac17, bsw, lsw, aw11, common, PL = mods
    rng = random.Random(2)
    names = ["a%d" % i for i in range(64)]
    policy = binary_tree(names, rng)
    pk, msk = ac17.setup(common.Rng(2))
    pts = [bytes([i % 251]) * (1 + i % 7) for i in range(512)]
    cts = ac17.cp_encrypt_batch(pk, policy, pts, PL.HumanPolicy, common.Rng(3))
    sk = ac17.cp_keygen(msk, names, common.Rng(4))
    assert ac17.cp_decrypt_batch(sk, cts) == pts
    sk_partial = ac17.cp_keygen(msk, names[:1], common.Rng(5))
    try:
        ok = ac17.cp_decrypt_batch(sk_partial, cts[:2]) == pts[:2]
    except ac17.RabeError:
        ok = False
    from rabe_b200.policy import Policy
    assert ok == Policy(policy, PL.HumanPolicy).satisfied(names[:1])


def test_config3_bsw_128_attribute_and_tree(mods):
    ac17, bsw, lsw, aw11, common, PL = mods
    names = ["a%d" % i for i in range(128)]
    policy = and_tree([names[8 * g:8 * g + 8] for g in range(16)])            # root AND over 16 x AND(8): nI = 128
    pk, msk = bsw.setup(common.Rng(3))
    ct = bsw.encrypt(pk, policy, PL.HumanPolicy, PLAINTEXT, common.Rng(31))
    assert len(ct.c_y) == 128
    sk = bsw.keygen(pk, msk, names, common.Rng(32))
    assert bsw.decrypt(sk, ct) == PLAINTEXT                                   # 257 pairings, one final exponentiation
    with pytest.raises(bsw.RabeError):
        bsw.decrypt(bsw.keygen(pk, msk, names[:-1], common.Rng(33)), ct)


def test_config4_lsw_256_attribute_policy(mods):
    ac17, bsw, lsw, aw11, common, PL = mods
    names = ["a%d" % i for i in range(256)]
    policy = and_tree([names[16 * g:16 * g + 16] for g in range(16)])         # root AND over 16 x AND(16)
    pk, msk = lsw.setup(common.Rng(4))
    sk = lsw.keygen(pk, msk, policy, PL.HumanPolicy, common.Rng(41))
    assert len(sk.dj) == 256
    ct = lsw.encrypt(pk, names, PLAINTEXT, common.Rng(42))
    assert lsw.decrypt(sk, ct) == PLAINTEXT                                   # 512 pairings in one product
    with pytest.raises(lsw.RabeError):
        lsw.decrypt(sk, lsw.encrypt(pk, names[1:], PLAINTEXT, common.Rng(43)))


def test_config5_aw11_8_authorities_x_32_attributes(mods):
    ac17, bsw, lsw, aw11, common, PL = mods
    rng = random.Random(5)
    gk = aw11.setup(common.Rng(5))
    auth_names = [["AUTH%dATTR%d" % (k, j) for j in range(32)] for k in range(8)]
    auths = [aw11.authgen(gk, names, common.Rng(50 + k)) for k, names in enumerate(auth_names)]
    flat = [n for names in auth_names for n in names]
    policy = binary_tree(flat, rng)                                           # binary AND/OR tree over all 256
    ct = aw11.encrypt(gk, [a[0] for a in auths], policy, PL.HumanPolicy, PLAINTEXT, common.Rng(58))
    assert len(ct.c) == 256
    sk = aw11.Aw11SecretKey("alice", [])
    for (pk, msk), names in zip(auths, auth_names):
        for n in names:
            aw11.add_to_attribute(gk, msk, n, sk)
    assert aw11.decrypt(gk, sk, ct) == PLAINTEXT
