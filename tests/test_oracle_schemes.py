"""The oracle's scheme restatements (oracle/schemes.py) pass rabe's own round-trip tests on the CPU:
matching keys recover the Gt message, non-matching keys fail (bsw/mod.rs:326-602, lsw/mod.rs:298-374,
aw11/mod.rs:398-561, ghw11/mod.rs:311-395).  Keeps the checker honest before the GPU is compared with it."""
import random

from oracle import policy as OP
from oracle import schemes as OS
from oracle.pyref import R


def draws(rng, n=400):
    return iter([rng.randrange(R) for _ in range(n)])


def test_ghw11_oracle_round_trips():
    rng = random.Random(7)
    pk, msk = OS.ghw11_setup(draws(rng))
    and10 = '{"name": "and", "children": [' + ", ".join('{"name": "attr%d"}' % n for n in range(1, 11)) + ']}'
    for text, attrs, bad in (('{"name": "or", "children": [{"name": "A"}, {"name": "B"}]}', ["D", "B"], ["C", "D"]),
                             ('{"name": "and", "children": [{"name": "attr0"}, {"name": "attr1"}]}', ["attr0", "attr1"], ["attr1"]),
                             (and10, ["attr%d" % n for n in range(1, 11)], ["attr201", "attr200"])):
        msg = OS.gt_random(rng.randrange(R))
        ct = OS.ghw11_encrypt(pk, text, OP.JSON, msg, draws(rng))
        tk, rk = OS.ghw11_tkgen(OS.ghw11_keygen(pk, msk, attrs, draws(rng)), draws(rng))
        assert OS.ghw11_decrypt_out(OS.ghw11_transform(ct, tk), rk) == msg
        bad_tk, _ = OS.ghw11_tkgen(OS.ghw11_keygen(pk, msk, bad, draws(rng)), draws(rng))
        assert OS.ghw11_transform(ct, bad_tk) is None
    assert OS.ghw11_keygen(pk, msk, [], draws(rng)) is None


def test_bsw_lsw_aw11_oracle_round_trips():
    rng = random.Random(8)
    msg = OS.gt_random(rng.randrange(R))
    pk, msk = OS.bsw_setup(draws(rng))
    ct = OS.bsw_encrypt(pk, '("A" and "B") or "C"', OP.HUMAN, msg, draws(rng))
    assert OS.bsw_decrypt(OS.bsw_keygen(pk, msk, ["B", "A"], draws(rng)), ct) == msg
    assert OS.bsw_decrypt(OS.bsw_keygen(pk, msk, ["A"], draws(rng)), ct) is None
    pk, msk = OS.lsw_setup(draws(rng))
    sk = OS.lsw_keygen(pk, msk, '("A" and "B") or "C"', OP.HUMAN, draws(rng))
    assert OS.lsw_decrypt(sk, OS.lsw_encrypt(pk, ["C", "Z"], msg, draws(rng))) == msg
    assert OS.lsw_decrypt(sk, OS.lsw_encrypt(pk, ["A", "Z"], msg, draws(rng))) is None
    gk = OS.aw11_setup(draws(rng))
    pk1, msk1 = OS.aw11_authgen(gk, ["A", "B"], draws(rng))
    ct = OS.aw11_encrypt(gk, [pk1], '"A" and "B"', OP.HUMAN, msg, draws(rng))
    assert OS.aw11_decrypt(gk, OS.aw11_keygen(gk, msk1, "bob", ["A", "B"]), ct) == msg
    assert OS.aw11_decrypt(gk, OS.aw11_keygen(gk, msk1, "bob", ["A"]), ct) is None
