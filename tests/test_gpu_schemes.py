"""BSW / LSW / AW11: the GPU scheme mirrors against the oracle's reference-sequence restatements
(oracle/schemes.py) on identical keys, policies, attributes and explicit randomness -- every group
element of every key and ciphertext compared byte for byte, plus rabe's own round-trip tests
(bsw/mod.rs:326-602, lsw/mod.rs:298-374, aw11/mod.rs:398-561)."""
import random

import pytest

import oracle
from oracle import policy as OP
from oracle import schemes as OS
from oracle.pyref import R

pytestmark = pytest.mark.gpu

PLAINTEXT = b"dance like no one's watching, encrypt like everyone is!"


def draws(rng, n=4000):
    return [rng.randrange(R) for _ in range(n)]


@pytest.fixture(scope="module")
def mods(engine):
    from rabe_b200.schemes import aw11, bsw, common, lsw
    from rabe_b200.policy import PolicyLanguage
    common.set_engine(engine)
    return bsw, lsw, aw11, common, PolicyLanguage


def fr(x):
    return int(x).to_bytes(32, "big")


def test_l0_fr_ops_and_adds(engine):
    from rb_testutil import u8
    rng = random.Random(40)
    a = [rng.randrange(R) for _ in range(33)]; b = [rng.randrange(R) for _ in range(33)]
    A, B = u8(b"".join(fr(x) for x in a)), u8(b"".join(fr(x) for x in b))
    assert engine.fr_op("add", A, B).tobytes() == b"".join(fr((x + y) % R) for x, y in zip(a, b))
    assert engine.fr_op("sub", A, B).tobytes() == b"".join(fr((x - y) % R) for x, y in zip(a, b))
    assert engine.fr_op("mul", A, B).tobytes() == b"".join(fr(x * y % R) for x, y in zip(a, b))
    assert engine.fr_op("mul", A, u8(fr(b[0]))).tobytes() == b"".join(fr(x * b[0] % R) for x in a)
    assert engine.fr_op("neg", A).tobytes() == b"".join(fr(-x % R) for x in a)
    assert engine.fr_op("inverse", A).tobytes() == b"".join(fr(pow(x, -1, R)) for x in a)
    g, h = oracle.g1_generator(), oracle.g2_generator()
    p = [oracle.g1_mul(g, fr(x)) for x in a[:5]]; q = [oracle.g2_mul(h, fr(x)) for x in a[:5]]
    assert engine.g1_add(u8(b"".join(p)), u8(b"".join(reversed(p)))).tobytes() == b"".join(oracle.g1_add(x, y) for x, y in zip(p, reversed(p)))
    assert engine.g2_add(u8(b"".join(q)), u8(q[0])).tobytes() == b"".join(oracle.g2_add(x, q[0]) for x in q)


def test_shares_and_coefficients(engine):
    from rabe_b200.policy import Policy, PolicyLanguage
    from rb_testutil import u8
    rng = random.Random(41)
    for text in ('{"name": "and", "children": [{"name": "A"}, {"name": "B"}, {"name": "C"}, {"name": "D"}]}',
                 '{"name": "or", "children": [{"name": "A"}, {"name": "and", "children": [{"name": "B"}, {"name": "or", "children": [{"name": "C"}, {"name": "and", "children": [{"name": "D"}, {"name": "E"}, {"name": "F"}]}]}]}]}',
                 '{"name": "A"}'):
        pol = Policy(text, PolicyLanguage.JsonPolicy)
        tree = OP.parse(text, OP.JSON)
        plan = engine.share_plan(pol)
        assert plan.n_coefs == OP.count_share_randomness(tree)
        B = 3
        secrets = [rng.randrange(R) for _ in range(B)]
        coefs = [[rng.randrange(R) for _ in range(plan.n_coefs)] for _ in range(B)]
        got = engine.shares(plan, u8(b"".join(fr(s) for s in secrets)), u8(b"".join(fr(c) for row in coefs for c in row))).tobytes()
        exp = b""
        for b in range(B):
            sh = OP.gen_shares_policy(secrets[b], tree, iter(coefs[b]))
            assert [l for l, _ in sh] == pol.leaf_labels()
            exp += b"".join(fr(v) for _, v in sh)
        assert got == exp
        oc = OP.calc_coefficients(tree)
        assert engine.policy_coefficients(pol, plan.n_leaves).tobytes() == b"".join(fr(c) for _, c in oc)


def test_bsw_parity_and_round_trips(mods):
    bsw, lsw, aw11, common, PL = mods
    rng = random.Random(42)
    d = draws(rng)
    opk, omsk = OS.bsw_setup(iter(d))
    pk, msk = bsw.setup(common.Rng(values=d))
    assert (pk.g1, pk.g2, pk.h, pk.f, pk.e_gg_alpha) == (opk["g1"], opk["g2"], opk["h"], opk["f"], opk["e_gg_alpha"])
    assert (int.from_bytes(msk.beta, "big"), msk.g2_alpha) == (omsk["beta"], omsk["g2_alpha"])
    msg = OS.gt_random(rng.randrange(R))
    cases = [  # or :326, and10 :359 (here 6-ary), or3 :441, and :472, and3 :527, or_and :556
        ('{"name": "or", "children": [{"name": "A"}, {"name": "B"}]}', PL.JsonPolicy, OP.JSON, ["B"]),
        ('{"name": "and", "children": [{"name": "A"}, {"name": "B"}, {"name": "C"}, {"name": "D"}, {"name": "E"}, {"name": "F"}]}', PL.JsonPolicy, OP.JSON, list("FEDCBA")),
        ('{"name": "or", "children": [{"name": "X"}, {"name": "Y"}, {"name": "A"}]}', PL.JsonPolicy, OP.JSON, ["A", "Q"]),
        ('"A" and "B"', PL.HumanPolicy, OP.HUMAN, ["A", "B"]),
        ('("A" and "B") or ("C" and ("D" or "E") and "F")', PL.HumanPolicy, OP.HUMAN, ["C", "E", "F", "A"]),
    ]
    for text, lang, olang, attrs in cases:
        d = draws(rng)
        oct_ = OS.bsw_encrypt(opk, text, olang, msg, iter(d))
        ct = bsw.encrypt(pk, text, lang, PLAINTEXT, common.Rng(values=d), _msg=msg)
        assert (ct.c, ct.c_p) == (oct_["c"], oct_["c_p"]), text
        assert [(x.string, x.g1, x.g2) for x in ct.c_y] == oct_["c_y"], text
        d = draws(rng)
        osk = OS.bsw_keygen(opk, omsk, attrs, iter(d))
        sk = bsw.keygen(pk, msk, attrs, common.Rng(values=d))
        assert sk.d == osk["d"] and [(x.string, x.g1, x.g2) for x in sk.d_j] == osk["d_j"]
        assert bsw.decrypt_gt(sk, ct) == OS.bsw_decrypt(osk, oct_) == msg
        assert bsw.decrypt(sk, ct) == PLAINTEXT
    # non-matching key: error like the reference; delegate_ab (:579)
    ct = bsw.encrypt(pk, '"A" and "B"', PL.HumanPolicy, PLAINTEXT, common.Rng(1))
    with pytest.raises(bsw.RabeError):
        bsw.decrypt(bsw.keygen(pk, msk, ["A"], common.Rng(2)), ct)
    assert bsw.keygen(pk, msk, [], common.Rng(2)) is None
    d = draws(rng)
    osk = OS.bsw_keygen(opk, omsk, ["A", "B", "C"], iter(d)); sk = bsw.keygen(pk, msk, ["A", "B", "C"], common.Rng(values=d))
    d = draws(rng)
    osk2 = OS.bsw_delegate(opk, osk, ["A", "B"], iter(d)); sk2 = bsw.delegate(pk, sk, ["A", "B"], common.Rng(values=d))
    assert sk2.d == osk2["d"] and [(x.string, x.g1, x.g2) for x in sk2.d_j] == osk2["d_j"]
    assert bsw.decrypt(sk2, ct) == PLAINTEXT
    assert bsw.delegate(pk, sk, ["A", "Z"], common.Rng(3)) is None


def test_lsw_parity_and_round_trips(mods):
    bsw, lsw, aw11, common, PL = mods
    rng = random.Random(43)
    d = draws(rng)
    opk, omsk = OS.lsw_setup(iter(d))
    pk, msk = lsw.setup(common.Rng(values=d))
    assert (pk.g1, pk.g2, pk.g1_b, pk.g1_b2, pk.h_b, pk.e_gg_alpha) == tuple(opk[k] for k in ("g1", "g2", "g1_b", "g1_b2", "h_b", "e_gg_alpha"))
    msg = OS.gt_random(rng.randrange(R))
    cases = [  # and :298, or :317, or_and :336
        ('{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}', PL.JsonPolicy, OP.JSON, ["A", "B"], True),
        ('{"name": "or", "children": [{"name": "A"}, {"name": "B"}]}', PL.JsonPolicy, OP.JSON, ["B", "Z"], True),
        ('("A" and "B" and "C") or ("D" and "E")', PL.HumanPolicy, OP.HUMAN, ["X", "A", "C", "B"], True),
        ('"A" and "B"', PL.HumanPolicy, OP.HUMAN, ["A", "C"], False),
    ]
    for text, lang, olang, attrs, ok in cases:
        d = draws(rng)
        osk = OS.lsw_keygen(opk, omsk, text, olang, iter(d))
        sk = lsw.keygen(pk, msk, text, lang, common.Rng(values=d))
        assert sk.dj == osk["dj"], text
        d = draws(rng)
        oct_ = OS.lsw_encrypt(opk, attrs, msg, iter(d))
        ct = lsw.encrypt(pk, attrs, PLAINTEXT, common.Rng(values=d), _msg=msg)
        assert (ct.e1, ct.e2, ct.ej) == (oct_["e1"], oct_["e2"], oct_["ej"]), text
        if ok:
            assert lsw.decrypt_gt(sk, ct) == OS.lsw_decrypt(osk, oct_) == msg
            assert lsw.decrypt(sk, ct) == PLAINTEXT
        else:
            assert OS.lsw_decrypt(osk, oct_) is None
            with pytest.raises(lsw.RabeError):
                lsw.decrypt(sk, ct)
    # negative attribute in the key policy: key parts match the reference (keygen :141-150)
    d = draws(rng)
    text = '"A" and "!B"'
    assert lsw.keygen(pk, msk, text, PL.HumanPolicy, common.Rng(values=d)).dj == OS.lsw_keygen(opk, omsk, text, OP.HUMAN, iter(d))["dj"]


def test_aw11_parity_and_round_trips(mods):
    bsw, lsw, aw11, common, PL = mods
    rng = random.Random(44)
    d = draws(rng)
    ogk = OS.aw11_setup(iter(d)); gk = aw11.setup(common.Rng(values=d))
    assert (gk.g1, gk.g2) == (ogk["g1"], ogk["g2"])
    auths = []
    for names in (["a", "B"], ["C", "D", "E"]):
        d = draws(rng)
        opk, omsk = OS.aw11_authgen(ogk, names, iter(d))
        pk, msk = aw11.authgen(gk, names, common.Rng(values=d))
        assert pk.attr == opk["attr"]
        assert [(n, int.from_bytes(a, "big"), int.from_bytes(y, "big")) for n, a, y in msk.attr] == omsk["attr"]
        auths.append((opk, omsk, pk, msk))
    msg = OS.gt_random(rng.randrange(R))
    cases = [  # and :398, or :437, or_and :478
        ('{"name": "and", "children": [{"name": "A"}, {"name": "C"}]}', PL.JsonPolicy, OP.JSON, [(0, "A"), (1, "C")], True),
        ('{"name": "or", "children": [{"name": "B"}, {"name": "E"}]}', PL.JsonPolicy, OP.JSON, [(1, "E")], True),
        ('("A" and "B") or ("C" and ("D" and "E"))', PL.HumanPolicy, OP.HUMAN, [(1, "C"), (1, "D"), (1, "E")], True),
        ('"A" and "C"', PL.HumanPolicy, OP.HUMAN, [(0, "A")], False),
    ]
    for text, lang, olang, held, ok in cases:
        d = draws(rng)
        oct_ = OS.aw11_encrypt(ogk, [a[0] for a in auths], text, olang, msg, iter(d))
        ct = aw11.encrypt(gk, [a[2] for a in auths], text, lang, PLAINTEXT, common.Rng(values=d), _msg=msg)
        assert ct.c_0 == oct_["c_0"] and ct.c == oct_["c"], text
        osk = {"gid": "bob", "attr": []}
        sk = aw11.Aw11SecretKey("bob", [])
        for ai, name in held:
            osk["attr"] += OS.aw11_keygen(ogk, auths[ai][1], "bob", [name])["attr"]
            aw11.add_to_attribute(gk, auths[ai][3], name, sk)
        assert sk.attr == osk["attr"]
        if ok:
            assert aw11.decrypt_gt(gk, sk, ct) == OS.aw11_decrypt(ogk, osk, oct_) == msg
            assert aw11.decrypt(gk, sk, ct) == PLAINTEXT
        else:
            assert OS.aw11_decrypt(ogk, osk, oct_) is None
            with pytest.raises(aw11.RabeError):
                aw11.decrypt(gk, sk, ct)
    assert aw11.authgen(gk, [], common.Rng(1)) is None
    with pytest.raises(aw11.RabeError):
        aw11.keygen(gk, auths[0][3], "", ["A"])


def test_ac17_kp_parity_and_round_trips(mods, engine):
    """ac17 kp_and :683, kp_or_and :701, kp_or :726 + element-wise parity with the oracle restatement."""
    bsw, lsw, aw11, common, PL = mods
    from rabe_b200.schemes import ac17
    rng = random.Random(45)
    setup_rnd = b"".join(fr(rng.randrange(R)) for _ in range(9))
    pkb, mskb = oracle.ac17_setup(setup_rnd)
    pk, msk = ac17.Ac17PublicKey.from_bytes(pkb), ac17.Ac17MasterKey.from_bytes(mskb)
    msg = OS.gt_random(rng.randrange(R))
    cases = [
        ('{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}', PL.JsonPolicy, OP.JSON, ["A", "B"], True),
        ('{"name": "or", "children": [{"name": "X"}, {"name": "and", "children": [{"name": "A"}, {"name": "B"}]}]}', PL.JsonPolicy, OP.JSON, ["A", "B", "C"], True),
        ('{"name": "or", "children": [{"name": "A"}, {"name": "B"}]}', PL.JsonPolicy, OP.JSON, ["B"], True),
        ('("A" and "B") and ("C" or ("D" and "E"))', PL.HumanPolicy, OP.HUMAN, ["A", "B", "D", "E"], True),
        ('"A" and "B"', PL.HumanPolicy, OP.HUMAN, ["A", "C"], False),
    ]
    for text, lang, olang, attrs, ok in cases:
        d = draws(rng)
        osk = OS.ac17_kp_keygen(mskb, text, olang, iter(d))
        sk = ac17.kp_keygen(msk, text, lang, common.Rng(values=d))
        assert sk.sk.k_0 == osk["k_0"] and sk.sk.k == osk["k"], text
        d = draws(rng)
        oct_ = OS.ac17_kp_encrypt(pkb, attrs, msg, iter(d))
        ct = ac17.kp_encrypt(pk, attrs, PLAINTEXT, common.Rng(values=d), _msg=msg)
        assert (ct.ct.c_0, ct.ct.c, ct.ct.c_p) == (oct_["c_0"], oct_["c"], oct_["c_p"]), text
        if ok:
            assert ac17.kp_decrypt_gt(sk, ct) == OS.ac17_kp_decrypt(osk, oct_) == msg
            assert ac17.kp_decrypt(sk, ct) == PLAINTEXT
        else:
            assert OS.ac17_kp_decrypt(osk, oct_) is None
            with pytest.raises(ac17.RabeError):
                ac17.kp_decrypt(sk, ct)


def test_fused_batch_entry_points_match_per_item_results(mods, engine):
    """rb_bsw_{encrypt,keygen,decrypt}_batch, rb_lsw_{keygen,decrypt}_batch and rb_aw11_encrypt_batch
    with B > 1: every item equals the oracle / the B = 1 call on the same randomness."""
    bsw, lsw, aw11, common, PL = mods
    from rabe_b200.policy import Policy, remove_index, sha3_hash_fr
    from rb_testutil import u8
    rng = random.Random(46)
    d = draws(rng)
    opk, omsk = OS.bsw_setup(iter(d)); pk, msk = bsw.setup(common.Rng(values=d))
    text, attrs = '("A" and "B") or ("C" and ("D" or "E") and "F")', ["C", "E", "F", "A"]
    B = 5
    msgs = [OS.gt_random(rng.randrange(R)) for _ in range(B)]
    d = draws(rng)
    it = iter(d)
    octs = [OS.bsw_encrypt(opk, text, OP.HUMAN, m, it) for m in msgs]
    cts = bsw.encrypt_batch(pk, text, PL.HumanPolicy, [PLAINTEXT] * B, common.Rng(values=d), _msgs=msgs)
    for ct, oct_ in zip(cts, octs):
        assert (ct.c, ct.c_p) == (oct_["c"], oct_["c_p"]) and [(x.string, x.g1, x.g2) for x in ct.c_y] == oct_["c_y"]
    # keygen: two keys in one call == two calls
    pkh = bsw._pk_handle(pk)
    hashes = u8(b"".join(sha3_hash_fr(a) for a in attrs))
    r = b"".join(fr(rng.randrange(R)) for _ in range(2)); rj = b"".join(fr(rng.randrange(R)) for _ in range(2 * len(attrs)))
    d2, g1s, g2s = [x.tobytes() for x in engine.bsw_keygen(pkh, u8(msk.beta), u8(msk.g2_alpha), hashes, u8(r), u8(rj))]
    n = len(attrs)
    for b in range(2):
        osk = OS.bsw_keygen(opk, omsk, attrs, iter([int.from_bytes(r[32 * b:32 * b + 32], "big")] +
                                                   [int.from_bytes(rj[32 * (b * n + i):32 * (b * n + i + 1)], "big") for i in range(n)]))
        assert d2[128 * b:128 * b + 128] == osk["d"]
        assert [(a, g1s[64 * (b * n + i):64 * (b * n + i + 1)], g2s[128 * (b * n + i):128 * (b * n + i + 1)]) for i, a in enumerate(attrs)] == osk["d_j"]
    # decrypt: the 5 ciphertexts in one call
    sk = bsw.keygen(pk, msk, attrs, common.Rng(9))
    pol = Policy(text, PL.HumanPolicy)
    ok, pruned = pol.prune(attrs)
    labels = pol.leaf_labels()
    z = engine.policy_coefficients(pol, len(labels)).tobytes()
    ct_names, sk_names = [x.string for x in cts[0].c_y], [x.string for x in sk.d_j]
    ct_idx = [ct_names.index(j) for _, j in pruned]; sk_idx = [sk_names.index(k) for k, _ in pruned]
    coeff = b"".join(z[32 * labels.index(j):32 * labels.index(j) + 32] for _, j in pruned)
    out = engine.bsw_decrypt(u8(sk.d), u8(b"".join(x.g1 for x in sk.d_j)), u8(b"".join(x.g2 for x in sk.d_j)),
                             u8(b"".join(c.c for c in cts)), u8(b"".join(c.c_p for c in cts)),
                             u8(b"".join(x.g1 for c in cts for x in c.c_y)), u8(b"".join(x.g2 for c in cts for x in c.c_y)),
                             ct_idx, sk_idx, u8(coeff)).tobytes()
    assert [out[384 * b:384 * b + 384] for b in range(B)] == msgs
    # LSW: 3 keys of one policy in one call; 4 ciphertexts in one decrypt call
    d = draws(rng)
    lpk, lmsk = lsw.setup(common.Rng(values=d))
    ltext, lattrs = '("A" and "B" and "C") or ("D" and "E")', ["X", "A", "C", "B"]
    lpol = Policy(ltext, PL.HumanPolicy)
    plan = engine.share_plan(lpol)
    llabels = lpol.leaf_labels()
    lh = u8(b"".join(sha3_hash_fr(remove_index(l)) for l in llabels))
    keys = [lsw.keygen(lpk, lmsk, ltext, PL.HumanPolicy, common.Rng(100 + i)) for i in range(3)]
    co, rn = b"", b""
    for i in range(3):
        rr = common.Rng(100 + i); co += rr.frs(plan.n_coefs); rn += rr.frs(plan.n_leaves)
    g1t, g2t = common.TABLES.get("g1", lpk.g1, 16), common.TABLES.get("g2", lpk.g2, 8)
    d1, d2b = [x.tobytes() for x in engine.lsw_keygen(g1t, g2t, plan, lh, u8(lmsk.alpha1), u8(lmsk.alpha2), u8(co), u8(rn))]
    nl = plan.n_leaves
    for i, k in enumerate(keys):
        assert [x[1] for x in k.dj] == [d1[64 * (i * nl + j):64 * (i * nl + j + 1)] for j in range(nl)]
        assert [x[2] for x in k.dj] == [d2b[128 * (i * nl + j):128 * (i * nl + j + 1)] for j in range(nl)]
    lmsgs = [OS.gt_random(rng.randrange(R)) for _ in range(4)]
    lcts = [lsw.encrypt(lpk, lattrs, PLAINTEXT, common.Rng(200 + i), _msg=m) for i, m in enumerate(lmsgs)]
    ok, lpruned = lpol.prune(lattrs)
    lz = engine.policy_coefficients(lpol, nl).tobytes()
    skn, ctn = [x[0] for x in keys[0].dj], [x[0] for x in lcts[0].ej]
    out = engine.lsw_decrypt(u8(b"".join(x[1] for x in keys[0].dj)), u8(b"".join(x[2] for x in keys[0].dj)), u8(b"".join(c.e1 for c in lcts)),
                             u8(b"".join(c.e2 for c in lcts)), u8(b"".join(x[1] for c in lcts for x in c.ej)),
                             [ctn.index(nm) for nm, _ in lpruned], [skn.index(nm) for nm, _ in lpruned],
                             u8(b"".join(lz[32 * llabels.index(l):32 * llabels.index(l) + 32] for _, l in lpruned))).tobytes()
    assert [out[384 * b:384 * b + 384] for b in range(4)] == lmsgs
    # AW11: 3 messages in one call == the per-message mirror (which is checked against the oracle above)
    gk = aw11.setup(common.Rng(7))
    pk1, _ = aw11.authgen(gk, ["A", "B"], common.Rng(71)); pk2, _ = aw11.authgen(gk, ["C"], common.Rng(72))
    atext = '("A" and "B") or "C"'
    apol = Policy(atext, PL.HumanPolicy)
    aplan = engine.share_plan(apol)
    alabels = apol.leaf_labels()
    rows = [aw11.find_pk_attr([pk1, pk2], remove_index(l.upper())) for l in alabels]
    amsgs = [OS.gt_random(rng.randrange(R)) for _ in range(3)]
    singles, S, SC, WC, RX = [], b"", b"", b"", b""
    for i, m in enumerate(amsgs):
        singles.append(aw11.encrypt(gk, [pk1, pk2], atext, PL.HumanPolicy, PLAINTEXT, common.Rng(300 + i), _msg=m))
        rr = common.Rng(300 + i); S += rr.fr(); SC += rr.frs(aplan.n_coefs); WC += rr.frs(aplan.n_coefs); RX += rr.frs(aplan.n_leaves)
    c0, c1, c2, c3 = [x.tobytes() for x in engine.aw11_encrypt(common.TABLES.get("g2", gk.g2, 8), common.TABLES.get("gt", aw11._e_gg(gk), 8), aplan,
                                                                u8(b"".join(a[1] for a in rows)), u8(b"".join(a[2] for a in rows)),
                                                                u8(S), u8(SC), u8(WC), u8(RX), u8(b"".join(amsgs)))]
    na = aplan.n_leaves
    for b, ct in enumerate(singles):
        assert ct.c_0 == c0[384 * b:384 * b + 384]
        assert [x[1] for x in ct.c] == [c1[384 * (b * na + j):384 * (b * na + j + 1)] for j in range(na)]
        assert [x[2] for x in ct.c] == [c2[128 * (b * na + j):128 * (b * na + j + 1)] for j in range(na)]
        assert [x[3] for x in ct.c] == [c3[128 * (b * na + j):128 * (b * na + j + 1)] for j in range(na)]
    # the same through per-attribute fixed-base tables (rb_aw11_pk_load / rb_aw11_encrypt_pk_batch), attributes stored in another order
    order = list(reversed(range(na)))                               # handle holds the attributes reversed; leaf i -> order.index(i)
    pkh = engine.aw11_pk_load(u8(b"".join(rows[j][1] for j in order)), u8(b"".join(rows[j][2] for j in order)))
    leaf_attr = [order.index(i) for i in range(na)]
    t0, t1, t2, t3 = [x.tobytes() for x in engine.aw11_encrypt_pk(common.TABLES.get("g2", gk.g2, 8), common.TABLES.get("gt", aw11._e_gg(gk), 8), aplan,
                                                                  pkh, leaf_attr, u8(S), u8(SC), u8(WC), u8(RX), u8(b"".join(amsgs)))]
    assert (t0, t1, t2, t3) == (c0, c1, c2, c3)



def test_lsw_encrypt_batch_matches_oracle_per_item(mods, engine):
    """rb_lsw_encrypt_batch (lsw/mod.rs:180-219) with B > 1: every member of every item equals the oracle on the
    same draws, including the reference's `sx[0]` quirk, for n = 1 (sx[0] = 0), n = 2 and n = 5."""
    bsw, lsw, aw11, common, PL = mods
    rng = random.Random(47)
    d = draws(rng)
    opk, omsk = OS.lsw_setup(iter(d)); pk, msk = lsw.setup(common.Rng(values=d))
    for attrs in (["A"], ["A", "B"], ["X", "A", "C", "B", "Q"]):
        B = 4
        msgs = [OS.gt_random(rng.randrange(R)) for _ in range(B)]
        d = draws(rng)
        it = iter(d)
        octs = [OS.lsw_encrypt(opk, attrs, m, it) for m in msgs]
        cts = lsw.encrypt_batch(pk, attrs, [PLAINTEXT] * B, common.Rng(values=d), _msgs=msgs)
        for ct, oct_ in zip(cts, octs):
            assert (ct.e1, ct.e2, ct.ej) == (oct_["e1"], oct_["e2"], oct_["ej"]), attrs
    with pytest.raises(lsw.RabeError):
        lsw.encrypt(pk, [], PLAINTEXT)


def test_ghw11_parity_and_round_trips(mods, engine):
    """GHW11 outsourced decryption (ghw11/mod.rs:92-305): every group element of pk / sk / tk / ct and the transformed
    Gt value against the oracle restatement, rabe's own tests (or :311, and2 :346, and10 :375), and B > 1 through the
    fused rb_ghw11_transform_batch."""
    bsw, lsw, aw11, common, PL = mods
    from rabe_b200.schemes import ghw11
    rng = random.Random(48)
    d = draws(rng)
    opk, omsk = OS.ghw11_setup(iter(d)); pk, msk = ghw11.setup(common.Rng(values=d))
    assert (pk.g1, pk.g2, pk.g1_a, pk.g2_a, pk.e_gg_alpha) == tuple(opk[k] for k in ("g1", "g2", "g1_a", "g2_a", "e_gg_alpha"))
    assert msk.g2_alpha == omsk["g2_alpha"]
    and10 = '{"name": "and", "children": [' + ", ".join('{"name": "attr%d"}' % n for n in range(1, 11)) + ']}'
    cases = [
        ('{"name": "or", "children": [{"name": "A"}, {"name": "B"}]}', PL.JsonPolicy, OP.JSON, ["D", "B"], ["C", "D"]),
        ('{"name": "and", "children": [{"name": "attr0"}, {"name": "attr1"}]}', PL.JsonPolicy, OP.JSON, ["attr0", "attr1"], ["attr0"]),
        (and10, PL.JsonPolicy, OP.JSON, ["attr%d" % n for n in range(1, 11)], ["attr201", "attr200"]),
        ('("A" and "B") or ("C" and ("D" or "E") and "F")', PL.HumanPolicy, OP.HUMAN, ["C", "E", "F", "A"], ["A", "C", "F"]),
    ]
    for text, lang, olang, attrs, bad_attrs in cases:
        msg = OS.gt_random(rng.randrange(R))
        d = draws(rng)
        oct_ = OS.ghw11_encrypt(opk, text, olang, msg, iter(d))
        ct = ghw11.encrypt(pk, text, lang, PLAINTEXT, common.Rng(values=d), _msg=msg)
        assert (ct.c, ct.c1, ct.ci_di) == (oct_["c"], oct_["c1"], oct_["ci_di"]), text
        d = draws(rng)
        osk = OS.ghw11_keygen(opk, omsk, attrs, iter(d)); sk = ghw11.keygen(pk, msk, attrs, common.Rng(values=d))
        assert (sk.k, sk.l, [(x.string, x.k_x) for x in sk.attr_key]) == (osk["k"], osk["l"], osk["attr_key"])
        d = draws(rng)
        otk, ork = OS.ghw11_tkgen(osk, iter(d)); tk, rk = ghw11.tkgen(sk, common.Rng(values=d))
        assert (tk.k_z, tk.l_z, [(x.string, x.k_x) for x in tk.attr_key_z]) == (otk["k_z"], otk["l_z"], otk["attr_key_z"])
        assert int.from_bytes(rk.z, "big") == ork["z"]
        opct = OS.ghw11_transform(oct_, otk); pct = ghw11.transform(ct, tk)
        assert (pct.c, pct.t) == (opct["c"], opct["t"]), text
        assert ghw11.decrypt_out_gt(pct, rk) == OS.ghw11_decrypt_out(opct, ork) == msg
        assert ghw11.decrypt_out(pct, rk, ct.data) == PLAINTEXT
        bad_tk, _ = ghw11.tkgen(ghw11.keygen(pk, msk, bad_attrs, common.Rng(5)), common.Rng(6))
        assert OS.ghw11_transform(oct_, OS.ghw11_tkgen(OS.ghw11_keygen(opk, omsk, bad_attrs, iter(draws(rng))), iter(draws(rng)))[0]) is None
        with pytest.raises(ghw11.RabeError):
            ghw11.transform(ct, bad_tk)
    assert ghw11.keygen(pk, msk, [], common.Rng(1)) is None
    # B = 5 ciphertexts of one policy through ONE fused transform call == the per-item oracle
    text, attrs = cases[3][0], cases[3][3]
    sk = ghw11.keygen(pk, msk, attrs, common.Rng(70)); tk, rk = ghw11.tkgen(sk, common.Rng(71))
    otk = {"k_z": tk.k_z, "l_z": tk.l_z, "attr_key_z": [(x.string, x.k_x) for x in tk.attr_key_z]}
    msgs = [OS.gt_random(rng.randrange(R)) for _ in range(5)]
    cts = [ghw11.encrypt(pk, text, PL.HumanPolicy, PLAINTEXT, common.Rng(80 + i), _msg=m) for i, m in enumerate(msgs)]
    pcts = ghw11.transform_batch(cts, tk)
    for ct, pct, m in zip(cts, pcts, msgs):
        o_ct = {"policy": (text, OP.HUMAN), "c": ct.c, "c1": ct.c1, "ci_di": ct.ci_di}
        assert pct.t == OS.ghw11_transform(o_ct, otk)["t"]
        assert ghw11.decrypt_out_gt(pct, rk) == m
