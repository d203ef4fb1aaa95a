"""GPU parity of the AC17 entry points against the reference-sequence oracle (same pk, policy,
attributes and explicit randomness) -- bit-exact on every canonical output byte."""
import random

import numpy as np
import pytest

import oracle
from oracle import policy as opol
import rb_testutil as util
from rb_testutil import fr, rand_fr, u8

pytestmark = pytest.mark.gpu


def _roundtrip(engine, policy, lang, key_attrs, B, seed, check_items=None):
    rng = random.Random(seed)
    setup_rnd = rand_fr(rng, 9)
    pk, msk = oracle.ac17_setup(setup_rnd)
    gpk, gmsk = engine.ac17_setup(u8(setup_rnd))
    assert gpk == pk and gmsk == msk
    tree = opol.parse(policy, lang)
    m, pi, n2 = opol.calculate_msp(tree)
    n1 = len(pi)
    h_row, h_col = util.ac17_hashes(pi, n2)
    pkh = engine.ac17_pk_load(u8(pk))
    msp = engine.msp_load(np.array(m, dtype=np.int8), u8(h_row), u8(h_col))
    s = rand_fr(rng, 2 * B)
    msgs = [util.gt_random(rng) for _ in range(B)]
    c0, c, cp = [x.tobytes() for x in engine.ac17_cp_encrypt(pkh, msp, u8(s), u8(b"".join(msgs)))]
    items = range(B) if check_items is None else check_items
    for b in items:
        e0, e1, e2 = oracle.ac17_cp_encrypt(pk, m, pi, s[64 * b:64 * b + 64], msgs[b])
        assert c0[384 * b:384 * (b + 1)] == e0, ("c_0", b)
        assert c[192 * n1 * b:192 * n1 * (b + 1)] == e1, ("c", b)
        assert cp[384 * b:384 * (b + 1)] == e2, ("c_p", b)
    # keygen
    n = len(key_attrs)
    h_attr, h_01 = util.ac17_attr_hashes(key_attrs)
    krnd = rand_fr(rng, n + 3)
    mskh = engine.ac17_msk_load(u8(msk))
    k0, k, kp = [x.tobytes() for x in engine.ac17_cp_keygen(mskh, u8(h_attr), u8(h_01), u8(krnd), n)]
    o0, o1, o2 = oracle.ac17_cp_keygen(msk, key_attrs, krnd)
    assert k0 == o0 and k == o1 and kp == o2
    # decrypt
    ok, pruned = opol.calc_pruned(key_attrs, tree)
    assert ok and opol.traverse_policy(key_attrs, tree)
    ct_idx, sk_idx = util.decrypt_lists(pruned, pi, key_attrs)
    out = engine.ac17_cp_decrypt(u8(k0), u8(k), u8(kp), u8(c0), u8(c), u8(cp), n1, ct_idx, sk_idx).tobytes()
    for b in range(B):
        assert out[384 * b:384 * (b + 1)] == msgs[b], ("msg", b)
    # same through a loaded key (fixed-argument Miller lines for k_0)
    skh = engine.ac17_sk_load(u8(k0), u8(k), u8(kp))
    out2 = engine.ac17_cp_decrypt_sk(skh, u8(c0), u8(c), u8(cp), n1, ct_idx, sk_idx).tobytes()
    assert out2 == out
    b = list(items)[0]
    ref = oracle.ac17_cp_decrypt([a for a, _ in pruned], pi, c0[384 * b:384 * (b + 1)], c[192 * n1 * b:192 * n1 * (b + 1)],
                                 cp[384 * b:384 * (b + 1)], key_attrs, k0, k, kp)
    assert ref == msgs[b]
    return pkh, msp


def test_config1_and4(engine):
    _roundtrip(engine, '("A" and "B") and ("C" and "D")', opol.HUMAN, ["A", "B", "C", "D"], B=5, seed=1)


def test_config1_json_twin(engine):
    pol = '{"name": "and", "children": [{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}, {"name": "and", "children": [{"name": "C"}, {"name": "D"}]}]}'
    _roundtrip(engine, pol, opol.JSON, ["A", "B", "C", "D"], B=2, seed=1)


def test_or_policies_and_superset_key(engine):
    _roundtrip(engine, '("A" or "X") and ("C" or "Y")', opol.HUMAN, ["Q", "C", "A", "B"], B=3, seed=7)
    _roundtrip(engine, '"A" or ("B" and "C")', opol.HUMAN, ["B", "C"], B=2, seed=8)
    _roundtrip(engine, '"A"', opol.HUMAN, ["A"], B=2, seed=9)


def test_duplicate_attribute_rows(engine):
    # the same attribute twice: the reference's name matching adds every matching row for every
    # list entry (ac17/mod.rs:404-413); decryption then does NOT recover msg -- parity, not success.
    policy = '("A" and "B") and ("A" and "C")'
    rng = random.Random(21)
    pk, msk = oracle.ac17_setup(rand_fr(rng, 9))
    tree = opol.parse(policy, opol.HUMAN)
    m, pi, n2 = opol.calculate_msp(tree)
    attrs = ["A", "B", "C"]
    h_row, h_col = util.ac17_hashes(pi, n2)
    s = rand_fr(rng, 2); msg = util.gt_random(rng)
    c0, c, cp = oracle.ac17_cp_encrypt(pk, m, pi, s, msg)
    krnd = rand_fr(rng, len(attrs) + 3)
    k0, k, kp = oracle.ac17_cp_keygen(msk, attrs, krnd)
    ok, pruned = opol.calc_pruned(attrs, tree)
    ct_idx, sk_idx = util.decrypt_lists(pruned, pi, attrs)
    ref = oracle.ac17_cp_decrypt([a for a, _ in pruned], pi, c0, c, cp, attrs, k0, k, kp)
    out = engine.ac17_cp_decrypt(u8(k0), u8(k), u8(kp), u8(c0), u8(c), u8(cp), len(pi), ct_idx, sk_idx).tobytes()
    assert out == ref


def test_config2_and64(engine):
    names = [f"a{i}" for i in range(64)]
    _roundtrip(engine, util.and_policy(names), opol.HUMAN, names, B=33, seed=2, check_items=[0, 17, 32])


def test_config2_random64(engine):
    names = [f"a{i}" for i in range(64)]
    pol = util.random_binary_policy(names, random.Random(2))
    _roundtrip(engine, pol, opol.HUMAN, names, B=4, seed=3, check_items=[0, 3])


def test_cp_encrypt_distinct_policies_per_item(engine):
    """policy_mode = distinct: every batch item has its own policy; the scalar
    tables are built on the device from SHA3 hashes computed on the device (rb_sha3_fr_batch,
    rb_msp_load_batch) and each item matches the oracle's cp_encrypt under its own policy."""
    import random
    import numpy as np
    import oracle
    from oracle import policy as opol
    from rb_testutil import R, fr, u8
    rng = random.Random(77)
    pk, msk = oracle.ac17_setup(b"".join(fr(rng.randrange(R)) for _ in range(9)))
    texts = ['("A" and "B") and ("C" and "D")', '("A" or "B") and ("C" or "D")', '("W" and "X") or ("Y" and "Z")',
             '"A" and ("B" and ("C" and "D"))', '(("A" or "B") or "C") or "D"', '("K" and "L") and ("M" or "N")']
    B = len(texts)
    msps = [opol.calculate_msp(opol.parse(t, opol.HUMAN)) for t in texts]
    n1 = 4
    n2 = max(c for _, _, c in msps)
    m = np.zeros((B, n1, n2), dtype=np.int8)
    row_strings, col_strings = [], []
    for b, (mm, pi, c) in enumerate(msps):
        assert len(pi) == n1
        m[b, :, :c] = np.array(mm, dtype=np.int8)
        row_strings += [("%s%d%d" % (nm, l, t)).encode() for nm in pi for l in range(3) for t in range(2)]
        col_strings += [("0%d%d%d" % (j + 1, l, t)).encode() for j in range(n2) for l in range(3) for t in range(2)]
    h_row, h_col = engine.sha3_fr(row_strings), engine.sha3_fr(col_strings)          # hashed on the device
    msp = engine.msp_load_batch(m, h_row, h_col)
    pkh = engine.ac17_pk_load(u8(pk))
    s = b"".join(fr(rng.randrange(R)) for _ in range(2 * B))
    e = oracle.pairing(oracle.g1_generator(), oracle.g2_generator())
    msgs = [oracle.gt_pow(e, fr(rng.randrange(R))) for _ in range(B)]
    c0, c, cp = [x.tobytes() for x in engine.ac17_cp_encrypt(pkh, msp, u8(s), u8(b"".join(msgs)))]
    for b, (mm, pi, _) in enumerate(msps):
        o0, oc, ocp = oracle.ac17_cp_encrypt(pk, mm, pi, s[64 * b:64 * b + 64], msgs[b])
        assert c0[384 * b:384 * b + 384] == o0 and cp[384 * b:384 * b + 384] == ocp, b
        assert c[192 * n1 * b:192 * n1 * (b + 1)] == oc, b
    # refold the same handle in place with the policies in another order (rb_msp_reload_batch), hashes from
    # the packed-input SHA3 entry point (rb_sha3_fr_batch_len): item b now encrypts under policy perm[b]
    perm = [3, 0, 5, 1, 4, 2]
    m2 = np.ascontiguousarray(m[perm])
    pack = lambda strs: (np.frombuffer(b"".join(strs), dtype=np.uint8), np.concatenate([[0], np.cumsum([len(x) for x in strs])]).astype(np.uint32))
    rows2 = [x for b in perm for x in row_strings[6 * n1 * b:6 * n1 * (b + 1)]]
    rd, ro = pack(rows2)
    cd, co = pack(col_strings)
    h_row2 = engine.sha3_fr_packed(rd, ro, len(rows2))
    h_col2 = engine.sha3_fr_packed(cd, co, len(col_strings))
    assert h_col2.tobytes() == h_col.tobytes()
    engine.msp_reload_batch(msp, m2, h_row2, h_col2)
    c0, c, cp = [x.tobytes() for x in engine.ac17_cp_encrypt(pkh, msp, u8(s), u8(b"".join(msgs)))]
    for b, src in enumerate(perm):
        mm, pi, _ = msps[src]
        o0, oc, ocp = oracle.ac17_cp_encrypt(pk, mm, pi, s[64 * b:64 * b + 64], msgs[b])
        assert c0[384 * b:384 * b + 384] == o0 and cp[384 * b:384 * b + 384] == ocp, b
        assert c[192 * n1 * b:192 * n1 * (b + 1)] == oc, b
    # the column labels do not depend on the policy: one shared [n2][3][2] table gives the same ciphertexts
    engine.msp_reload_batch(msp, m2, h_row2, h_col2[:192 * n2], h_col_shared=True)
    c0s, cs, cps = [x.tobytes() for x in engine.ac17_cp_encrypt(pkh, msp, u8(s), u8(b"".join(msgs)))]
    assert (c0s, cs, cps) == (c0, c, cp)
    # a per-item handle refuses a batch of another size
    from rabe_b200._lib import RabeB200Error
    import pytest as _pt
    with _pt.raises(RabeB200Error):
        engine.ac17_cp_encrypt(pkh, msp, u8(s[:64]), u8(msgs[0]))
