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
