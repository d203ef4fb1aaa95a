"""BASELINE.json configurations 2-5 at their full attribute counts: size-independent properties
(encrypt -> decrypt round trips, wrong-key failure) and -- one item per scheme -- EVERY group element
of the key, the ciphertext and the decrypted Gt value compared with the reference-sequence oracle
(oracle/schemes.py restating bsw/mod.rs:217-318, lsw/mod.rs:121-290, aw11/mod.rs:241-350) at the
full size: degree-7/15 share polynomials, 8/16-point Lagrange sets, the four-pair accumulator tail
when nI % 4 != 0, the e2 collapse over 256 points.  (B = 4096 of config 2 is oracle-sampled in
test_gpu_bench_config.py.)"""
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


# ---------------------------------------------------------------------------------------------
# one oracle-compared item per scheme at the BASELINE sizes (every group element, not just the round trip)
def _draws(rng, n):
    from oracle.pyref import R
    return [rng.randrange(R) for _ in range(n)]


def test_config3_bsw_128_every_element_vs_oracle(mods):
    from oracle import policy as OP, schemes as OS
    from oracle.pyref import R
    ac17, bsw, lsw, aw11, common, PL = mods
    rng = random.Random(303)
    names = ["a%d" % i for i in range(128)]
    policy = and_tree([names[8 * g:8 * g + 8] for g in range(16)])            # 16 x AND(8): degree-7 and degree-15 polynomials
    d = _draws(rng, 8)
    opk, omsk = OS.bsw_setup(iter(d)); pk, msk = bsw.setup(common.Rng(values=d))
    assert (pk.g1, pk.g2, pk.h, pk.f, pk.e_gg_alpha) == (opk["g1"], opk["g2"], opk["h"], opk["f"], opk["e_gg_alpha"])
    msg = OS.gt_random(rng.randrange(R))
    d = _draws(rng, 400)
    oct_ = OS.bsw_encrypt(opk, policy, OP.HUMAN, msg, iter(d))
    ct = bsw.encrypt(pk, policy, PL.HumanPolicy, PLAINTEXT, common.Rng(values=d), _msg=msg)
    assert (ct.c, ct.c_p) == (oct_["c"], oct_["c_p"])
    assert [(x.string, x.g1, x.g2) for x in ct.c_y] == oct_["c_y"] and len(ct.c_y) == 128
    d = _draws(rng, 200)
    osk = OS.bsw_keygen(opk, omsk, names, iter(d)); sk = bsw.keygen(pk, msk, names, common.Rng(values=d))
    assert sk.d == osk["d"] and [(x.string, x.g1, x.g2) for x in sk.d_j] == osk["d_j"]
    assert bsw.decrypt_gt(sk, ct) == OS.bsw_decrypt(osk, oct_) == msg         # 257 reference pairings vs one fused product


@pytest.mark.parametrize("last_group", [16, 14])                               # 256 leaves (nI % 4 == 0) and 254 (nI % 4 == 2)
def test_config4_lsw_256_every_element_vs_oracle(mods, last_group):
    from oracle import policy as OP, schemes as OS
    from oracle.pyref import R
    ac17, bsw, lsw, aw11, common, PL = mods
    rng = random.Random(404 + last_group)
    names = ["a%d" % i for i in range(240 + last_group)]
    policy = and_tree([names[16 * g:16 * g + 16] for g in range(16)])
    d = _draws(rng, 16)
    opk, omsk = OS.lsw_setup(iter(d)); pk, msk = lsw.setup(common.Rng(values=d))
    assert (pk.g1, pk.g2, pk.g1_b, pk.g1_b2, pk.h_b, pk.e_gg_alpha) == tuple(opk[k] for k in ("g1", "g2", "g1_b", "g1_b2", "h_b", "e_gg_alpha"))
    d = _draws(rng, 800)
    osk = OS.lsw_keygen(opk, omsk, policy, OP.HUMAN, iter(d)); sk = lsw.keygen(pk, msk, policy, PL.HumanPolicy, common.Rng(values=d))
    assert sk.dj == osk["dj"] and len(sk.dj) == len(names)
    msg = OS.gt_random(rng.randrange(R))
    d = _draws(rng, 400)
    oct_ = OS.lsw_encrypt(opk, names, msg, iter(d)); ct = lsw.encrypt(pk, names, PLAINTEXT, common.Rng(values=d), _msg=msg)
    assert (ct.e1, ct.e2, ct.ej) == (oct_["e1"], oct_["e2"], oct_["ej"])
    assert lsw.decrypt_gt(sk, ct) == OS.lsw_decrypt(osk, oct_) == msg


def test_config5_aw11_256_every_element_vs_oracle(mods):
    from oracle import policy as OP, schemes as OS
    from oracle.pyref import R
    ac17, bsw, lsw, aw11, common, PL = mods
    rng = random.Random(505)
    d = _draws(rng, 4)
    ogk = OS.aw11_setup(iter(d)); gk = aw11.setup(common.Rng(values=d))
    auth_names = [["AUTH%dATTR%d" % (k, j) for j in range(32)] for k in range(8)]
    auths = []
    for names in auth_names:
        d = _draws(rng, 80)
        opk, omsk = OS.aw11_authgen(ogk, names, iter(d)); pk, msk = aw11.authgen(gk, names, common.Rng(values=d))
        assert pk.attr == opk["attr"]
        auths.append((opk, omsk, pk, msk))
    flat = [n for names in auth_names for n in names]
    policy = binary_tree(flat, random.Random(5), p_and=0.85)                   # mostly AND: a large pruned set
    msg = OS.gt_random(rng.randrange(R))
    d = _draws(rng, 1200)
    oct_ = OS.aw11_encrypt(ogk, [a[0] for a in auths], policy, OP.HUMAN, msg, iter(d))
    ct = aw11.encrypt(gk, [a[2] for a in auths], policy, PL.HumanPolicy, PLAINTEXT, common.Rng(values=d), _msg=msg)
    assert ct.c_0 == oct_["c_0"] and ct.c == oct_["c"] and len(ct.c) == 256
    osk = {"gid": "bob", "attr": []}
    sk = aw11.Aw11SecretKey("bob", [])
    for (opk, omsk, pk, msk), names in zip(auths, auth_names):
        osk["attr"] += OS.aw11_keygen(ogk, omsk, "bob", names)["attr"]
        for n in names:
            aw11.add_to_attribute(gk, msk, n, sk)
    assert sk.attr == osk["attr"]
    assert aw11.decrypt_gt(gk, sk, ct) == OS.aw11_decrypt(ogk, osk, oct_) == msg
