"""Oracle parity on the configuration bench.py actually measures ("what's weak" 1-3):
the public key loaded with rb_ac17_pk_load_ex(26, 16, 16) (and 24/16/16), B = 4096 items of the
64-attribute policies, device-resident buffers, a loaded secret key -- and a seeded SAMPLE of the
batch compared element by element (c_0, c, c_p, decrypted Gt) with the reference-sequence oracle
(oracle/ac17.cpp restating ac17/mod.rs:274-430).  The full batch is additionally checked through
the round-trip property decrypt(encrypt(msg)) == msg.
"""
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle
from oracle import policy as opol
import rb_testutil as util
from rb_testutil import fr, rand_fr, u8

pytestmark = pytest.mark.gpu

B = 4096
N_SAMPLE = 16


def _bench_setup(engine, seed):
    rng = random.Random(seed)
    setup_rnd = rand_fr(rng, 9)
    pk, msk = oracle.ac17_setup(setup_rnd)
    names = [f"a{i}" for i in range(64)]
    krnd = rand_fr(rng, 64 + 3)
    k0, k, kp = oracle.ac17_cp_keygen(msk, names, krnd)
    return rng, pk, msk, names, (k0, k, kp)


def _check_policy(engine, pkh, pk, key, names, policy, rng, torch):
    dev = torch.device("cuda", 0)
    tree = opol.parse(policy, opol.HUMAN)
    m, pi, n2 = opol.calculate_msp(tree)
    n1 = len(pi)
    h_row, h_col = util.ac17_hashes(pi, n2)
    msp = engine.msp_load(np.array(m, dtype=np.int8), u8(h_row), u8(h_col))
    s = rand_fr(rng, 2 * B)
    # B distinct Gt messages: e_gh_ka[0]^rho through the GPU table, the sampled ones re-derived by the oracle below
    rho = rand_fr(rng, B)
    gt_tab = engine.gt_table(u8(pk[448:832]), 8)
    msg = engine.gt_pow_fixed(gt_tab, u8(rho)).tobytes()
    to_dev = lambda b: torch.from_numpy(u8(b)).to(dev)
    s_d, msg_d = to_dev(s), to_dev(msg)
    c0_d, c_d, cp_d = engine.ac17_cp_encrypt(pkh, msp, s_d, msg_d)
    k0, k, kp = key
    skh = engine.ac17_sk_load(u8(k0), u8(k), u8(kp))
    ok, pruned = opol.calc_pruned(names, tree)
    assert ok
    ct_idx, sk_idx = util.decrypt_lists(pruned, pi, names)
    ci_d = torch.from_numpy(np.array(ct_idx, dtype=np.uint32).view(np.int32)).to(dev)
    si_d = torch.from_numpy(np.array(sk_idx, dtype=np.uint32).view(np.int32)).to(dev)
    out_d = engine.ac17_cp_decrypt_sk(skh, c0_d, c_d, cp_d, n1, ci_d, si_d)       # rb_ac17_cp_decrypt_sk_batch, device buffers
    engine.status()
    c0, c, cp, out = [x.cpu().numpy().tobytes() for x in (c0_d, c_d, cp_d, out_d)]
    assert out == msg, "round trip over the whole batch"
    sample = sorted(random.Random(1234).sample(range(B), N_SAMPLE - 2) + [0, B - 1])
    plist = [a for a, _ in pruned]

    def one(b):
        m_b = oracle.gt_pow(pk[448:832], rho[32 * b:32 * b + 32])
        e0, e1, e2 = oracle.ac17_cp_encrypt(pk, m, pi, s[64 * b:64 * b + 64], m_b)
        dec = oracle.ac17_cp_decrypt(plist, pi, e0, e1, e2, names, k0, k, kp)
        return b, m_b, e0, e1, e2, dec

    with ThreadPoolExecutor(max_workers=8) as ex:          # ctypes releases the GIL
        for b, m_b, e0, e1, e2, dec in ex.map(one, sample):
            assert msg[384 * b:384 * (b + 1)] == m_b, ("msg", b)
            assert c0[384 * b:384 * (b + 1)] == e0, ("c_0", b)
            assert c[192 * n1 * b:192 * n1 * (b + 1)] == e1, ("c", b)
            assert cp[384 * b:384 * (b + 1)] == e2, ("c_p", b)
            assert out[384 * b:384 * (b + 1)] == dec == m_b, ("decrypt", b)


@pytest.mark.parametrize("windows", [(26, 16, 16), (24, 16, 16)])
def test_bench_configuration_against_oracle(engine, windows):
    import torch
    rng, pk, msk, names, key = _bench_setup(engine, seed=2)
    free, _total = torch.cuda.mem_get_info(0)
    need = (10 if windows[0] == 26 else 11) * (32 << windows[0]) + (6 << 30)          # signed-digit table: 2^(w-1) entries of 64 B per window
    if free < need:
        pytest.skip("not enough free HBM for a %d-bit G1 table" % windows[0])
    pkh = engine.ac17_pk_load(u8(pk), *windows)
    try:
        _check_policy(engine, pkh, pk, key, names, util.and_policy(names), rng, torch)                       # headline: all-AND, n1 = n2 = nI = 64
        _check_policy(engine, pkh, pk, key, names, util.random_binary_policy(names, random.Random(2)), rng, torch)   # random AND/OR tree
    finally:
        pkh.close()
        torch.cuda.empty_cache()
