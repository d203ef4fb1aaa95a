"""Committed golden vectors (tests/golden/, written by tools/gen_golden.py):
bn254_kat.json  -- from the independent pure-Python statement oracle/pyref.py, plus the public
                   EIP-196 2*G1 vector;
ac17_config1.json -- BASELINE.json config 1 through the C++ oracle with seeded randomness.
CPU: the C++ oracle reproduces both.  GPU: the CUDA path reproduces both through the C ABI."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from oracle import policy as opol

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "bn254_kat.json")))
CFG = json.load(open(os.path.join(HERE, "golden", "ac17_config1.json")))
G1, G2 = oracle.g1_generator(), oracle.g2_generator()
hx = bytes.fromhex
u8 = lambda b: np.frombuffer(bytes(b), dtype=np.uint8).copy()
sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


def test_oracle_matches_pyref_vectors():
    assert oracle.g1_mul(G1, (2).to_bytes(32, "big")) == hx(KAT["public_known_answer"]["out"])
    for v in KAT["g1_mul"]:
        assert oracle.g1_mul(G1, hx(v["k"])) == hx(v["out"])
    for v in KAT["g2_mul"]:
        assert oracle.g2_mul(G2, hx(v["k"])) == hx(v["out"])
    for v in KAT["pairing_lineage"]:
        assert oracle.g1_mul(G1, hx(v["a"])) == hx(v["p"]) and oracle.g2_mul(G2, hx(v["b"])) == hx(v["q"])
        assert oracle.pairing(hx(v["p"]), hx(v["q"])) == hx(v["e"])
    v = KAT["gt_pow"]
    assert oracle.gt_pow(hx(v["base"]), hx(v["k"])) == hx(v["out"])
    assert oracle.sha3_fr("A00") == hx(KAT["sha3_fr"]["A00"])


def _oracle_config1():
    pk, msk = oracle.ac17_setup(hx(CFG["setup_rnd"]))
    tree = opol.parse(CFG["policy"], opol.HUMAN)
    m, pi, n2 = opol.calculate_msp(tree)
    assert (m, pi, n2) == (CFG["msp"]["m"], CFG["msp"]["pi"], CFG["msp"]["n2"])
    c0, c, cp = oracle.ac17_cp_encrypt(pk, m, pi, hx(CFG["s"]), hx(CFG["msg"]))
    k0, k, kp = oracle.ac17_cp_keygen(msk, CFG["attrs"], hx(CFG["keygen_rnd"]))
    return pk, msk, c0, c, cp, k0, k, kp, pi, tree


def test_oracle_reproduces_ac17_config1():
    pk, msk, c0, c, cp, k0, k, kp, pi, tree = _oracle_config1()
    got = {"pk": sha(pk), "msk": sha(msk), "c_0": sha(c0), "c": sha(c), "c_p": sha(cp), "k_0": sha(k0), "k": sha(k), "k_p": sha(kp)}
    assert got == CFG["sha256"]
    ok, pruned = opol.calc_pruned(CFG["attrs"], tree)
    assert oracle.ac17_cp_decrypt([a for a, _ in pruned], pi, c0, c, cp, CFG["attrs"], k0, k, kp) == hx(CFG["msg"])


@pytest.mark.gpu
def test_gpu_matches_pyref_vectors(engine):
    ks = b"".join(hx(v["k"]) for v in KAT["g1_mul"])
    assert engine.g1_mul_var(u8(G1 * len(KAT["g1_mul"])), u8(ks)).tobytes() == b"".join(hx(v["out"]) for v in KAT["g1_mul"])
    tab = engine.g1_table(u8(G1), 16)
    assert engine.g1_mul_fixed(tab, u8(ks)).tobytes() == b"".join(hx(v["out"]) for v in KAT["g1_mul"])
    ks2 = b"".join(hx(v["k"]) for v in KAT["g2_mul"])
    assert engine.g2_mul_var(u8(G2 * len(KAT["g2_mul"])), u8(ks2)).tobytes() == b"".join(hx(v["out"]) for v in KAT["g2_mul"])
    P = b"".join(hx(v["p"]) for v in KAT["pairing_lineage"]); Q = b"".join(hx(v["q"]) for v in KAT["pairing_lineage"])
    assert engine.pairing(u8(P), u8(Q)).tobytes() == b"".join(hx(v["e"]) for v in KAT["pairing_lineage"])
    v = KAT["gt_pow"]
    assert engine.gt_pow_var(u8(hx(v["base"])), u8(hx(v["k"]))).tobytes() == hx(v["out"])


@pytest.mark.gpu
def test_gpu_reproduces_ac17_config1(engine):
    import ctypes
    from rabe_b200.engine import _Handle
    from rabe_b200.policy import Policy, PolicyLanguage
    pk, msk = engine.ac17_setup(u8(hx(CFG["setup_rnd"])))
    assert (sha(pk), sha(msk)) == (CFG["sha256"]["pk"], CFG["sha256"]["msk"])
    pol = Policy(CFG["policy"], PolicyLanguage.HumanPolicy)
    msp = engine.ac17_msp_from_policy(pol)
    pkh = engine.ac17_pk_load(u8(pk))
    c0, c, cp = [x.tobytes() for x in engine.ac17_cp_encrypt(pkh, msp, u8(hx(CFG["s"])), u8(hx(CFG["msg"])))]
    assert (sha(c0), sha(c), sha(cp)) == (CFG["sha256"]["c_0"], CFG["sha256"]["c"], CFG["sha256"]["c_p"])
    # keys + decrypt through the oracle-independent golden digests
    opk, omsk, oc0, oc, ocp, k0, k, kp, pi, tree = _oracle_config1()
    assert (sha(k0), sha(k), sha(kp)) == (CFG["sha256"]["k_0"], CFG["sha256"]["k"], CFG["sha256"]["k_p"])
    ok, pruned = opol.calc_pruned(CFG["attrs"], tree)
    ct_idx = np.array([pi.index(a) for a, _ in pruned], dtype=np.uint32)
    sk_idx = np.array([CFG["attrs"].index(a) for a, _ in pruned], dtype=np.uint32)
    out = engine.ac17_cp_decrypt(u8(k0), u8(k), u8(kp), u8(c0), u8(c), u8(cp), len(pi), ct_idx, sk_idx).tobytes()
    assert out == hx(CFG["msg"])
