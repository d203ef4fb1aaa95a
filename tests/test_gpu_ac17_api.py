"""rabe's own AC17 CP tests (ac17/mod.rs:759-809) replayed against the API mirror
rabe_b200.schemes.ac17 -- same policies, attribute sets, plaintext and expectations."""
import pytest

pytestmark = pytest.mark.gpu

PLAINTEXT = b"dance like no one's watching, encrypt like everyone is!"


@pytest.fixture(scope="module")
def ac17(engine):
    from rabe_b200.schemes import ac17 as mod
    mod.set_engine(engine)
    return mod


def test_cp_and(ac17):                      # ac17/mod.rs:760
    from rabe_b200.policy import PolicyLanguage
    pk, msk = ac17.setup(ac17.Rng(1))
    policy = '{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}'
    ct = ac17.cp_encrypt(pk, policy, PLAINTEXT, PolicyLanguage.JsonPolicy, ac17.Rng(2))
    sk = ac17.cp_keygen(msk, ["A", "B"], ac17.Rng(3))
    assert ac17.cp_decrypt(sk, ct) == PLAINTEXT
    assert [n for n, _ in ct.ct.c] == ["A", "B"] and len(ct.ct.c_0) == 3 and len(ct.ct.c_p) == 384
    assert ct.ct.ct[:12] != b"\0" * 12 and len(ct.ct.ct) == 12 + len(PLAINTEXT) + 16
    # non-matching key: error, not garbage
    sk_bad = ac17.cp_keygen(msk, ["A", "C"], ac17.Rng(4))
    with pytest.raises(ac17.RabeError):
        ac17.cp_decrypt(sk_bad, ct)
    with pytest.raises(ac17.RabeError):
        ac17.cp_keygen(msk, [], ac17.Rng(5))            # "empty attributes!" ac17/mod.rs:197


def test_cp_or(ac17):                       # ac17/mod.rs:777
    from rabe_b200.policy import PolicyLanguage
    pk, msk = ac17.setup(ac17.Rng(11))
    policy = '{"name": "or", "children": [{"name": "A"}, {"name": "B"}, {"name": "C"}]}'
    ct = ac17.cp_encrypt(pk, policy, PLAINTEXT, PolicyLanguage.JsonPolicy, ac17.Rng(12))
    sk = ac17.cp_keygen(msk, ["A"], ac17.Rng(13))
    assert ac17.cp_decrypt(sk, ct) == PLAINTEXT


def test_cp_or_and_and(ac17):               # ac17/mod.rs:795
    from rabe_b200.policy import PolicyLanguage
    pk, msk = ac17.setup(ac17.Rng(21))
    policy = '{"name": "or", "children": [{"name": "and", "children": [{"name": "A"}, {"name": "B"}]}, {"name": "and", "children": [{"name": "C"}, {"name": "D"}]}]}'
    ct = ac17.cp_encrypt(pk, policy, PLAINTEXT, PolicyLanguage.JsonPolicy, ac17.Rng(22))
    sk = ac17.cp_keygen(msk, ["A", "B", "C", "D"], ac17.Rng(23))
    assert ac17.cp_decrypt(sk, ct) == PLAINTEXT


def test_human_batch_and_seeded_determinism(ac17):
    from rabe_b200.policy import PolicyLanguage
    pk, msk = ac17.setup(ac17.Rng(31))
    pk2, msk2 = ac17.setup(ac17.Rng(31))
    assert pk == pk2 and msk == msk2
    policy = '("A" and "B") and ("C" or "D")'
    pts = [b"m%d" % i * (i + 1) for i in range(9)]
    cts = ac17.cp_encrypt_batch(pk, policy, pts, PolicyLanguage.HumanPolicy, ac17.Rng(32))
    sk = ac17.cp_keygen(msk, ["D", "B", "A"], ac17.Rng(33))
    assert ac17.cp_decrypt_batch(sk, cts) == pts
    tampered = cts[0]
    tampered.ct.c_p = cts[1].ct.c_p
    with pytest.raises(ac17.RabeError):
        ac17.cp_decrypt(sk, tampered)                   # AES-GCM tag check fails
