"""Asynchronous host-buffer mode (rb_ctx_set_async): calls with HOST buffers enqueue their copies and kernels and
return; results and the sticky status arrive with rb_ctx_sync() / rb_ctx_status().  One host thread then drives
several contexts, ordered by CUDA events on rb_ctx_get_stream()."""
import random

import numpy as np
import pytest

import oracle
from oracle import pyref as r
from rb_testutil import fr, u8

pytestmark = pytest.mark.gpu


def test_async_host_calls_complete_at_sync_and_report_errors_at_status():
    import torch
    from rabe_b200._lib import RabeB200Error, RB_ENOTMEMBER
    from rabe_b200.engine import Engine
    eng = Engine(0)
    rng = random.Random(31)
    g1 = oracle.g1_generator()
    tab = eng.g1_table(u8(g1), 12)
    n = 2000
    ks = [rng.randrange(r.R) for _ in range(n)]
    k_pin = torch.from_numpy(u8(b"".join(fr(k) for k in ks))).pin_memory()
    out_pin = torch.zeros(64 * n, dtype=torch.uint8).pin_memory()
    want = eng.g1_mul_fixed(tab, k_pin.numpy()).tobytes()                       # synchronous reference
    assert want[:64] == oracle.g1_mul(g1, fr(ks[0])) and want[-64:] == oracle.g1_mul(g1, fr(ks[-1]))
    eng.set_async(True)
    try:
        eng._call("rb_g1_mul_fixed_batch", tab, k_pin.numpy(), n, out_pin.numpy())   # returns after enqueueing
        eng.sync()
        assert out_pin.numpy().tobytes() == want
        # an invalid scalar (>= r): the call itself returns RB_OK, the status call reports it once
        bad = torch.from_numpy(u8(b"\xff" * 32 + fr(5))).pin_memory()
        out2 = torch.zeros(128, dtype=torch.uint8).pin_memory()
        eng._call("rb_g1_mul_fixed_batch", tab, bad.numpy(), 2, out2.numpy())
        with pytest.raises(RabeB200Error) as ei:
            eng.status()
        assert ei.value.status == RB_ENOTMEMBER
        eng.status()                                                               # cleared
        # handle-building calls still complete before returning
        tab2 = eng.g1_table(u8(oracle.g1_mul(g1, fr(7))), 8)
        eng._call("rb_g1_mul_fixed_batch", tab2, k_pin.numpy()[:64], 2, out2.numpy())
        eng.status()
        assert out2.numpy().tobytes() == oracle.g1_mul(g1, fr(7 * ks[0])) + oracle.g1_mul(g1, fr(7 * ks[1]))
    finally:
        eng.set_async(False)
        eng.close()


def test_one_host_thread_drives_two_contexts_through_host_buffers():
    """encrypt on context A (ciphertext lands in pinned host memory), decrypt on context B reading that host buffer,
    ordered only by a CUDA event between the two contexts' streams -- the pattern of bench.py's e2e pipeline."""
    import torch
    from rabe_b200.engine import Engine
    from oracle import policy as opol
    import rb_testutil as util
    A, Bc = Engine(0), Engine(0)
    rng = random.Random(32)
    pk, msk = oracle.ac17_setup(util.rand_fr(rng, 9))
    names = ["A", "B", "C"]
    tree = opol.parse('("A" and "B") or "C"', opol.HUMAN)
    m, pi, n2 = opol.calculate_msp(tree)
    h_row, h_col = util.ac17_hashes(pi, n2)
    k0, k, kp = oracle.ac17_cp_keygen(msk, names, util.rand_fr(rng, len(names) + 3))
    ok, pruned = opol.calc_pruned(names, tree)
    ct_idx, sk_idx = util.decrypt_lists(pruned, pi, names)
    pkh = A.ac17_pk_load(u8(pk)); msp = A.msp_load(np.array(m, dtype=np.int8), u8(h_row), u8(h_col))
    skh = Bc.ac17_sk_load(u8(k0), u8(k), u8(kp))
    n, n1 = 64, len(pi)
    pin = lambda b: torch.from_numpy(u8(b)).pin_memory()
    zeros = lambda nb: torch.zeros(nb, dtype=torch.uint8).pin_memory()
    sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(sA):
        A.use_torch_stream()
    with torch.cuda.stream(sB):
        Bc.use_torch_stream()
    A.set_async(True); Bc.set_async(True)
    rounds = []
    for it in range(4):
        s = pin(util.rand_fr(rng, 2 * n)); msgs = [util.gt_random(rng) for _ in range(n)]; msg = pin(b"".join(msgs))
        c0, c, cp, out = zeros(384 * n), zeros(192 * n1 * n), zeros(384 * n), zeros(384 * n)
        with torch.cuda.stream(sA):
            A.ac17_cp_encrypt(pkh, msp, s.numpy(), msg.numpy(), out=(c0.numpy(), c.numpy(), cp.numpy()))
            ev = torch.cuda.Event(); ev.record(sA)
        sB.wait_event(ev)
        with torch.cuda.stream(sB):
            Bc.ac17_cp_decrypt_sk(skh, c0.numpy(), c.numpy(), cp.numpy(), n1, ct_idx, sk_idx, out=out.numpy())
        rounds.append((s, msg, c0, c, cp, out, msgs))
    A.status(); Bc.status()
    for s, msg, c0, c, cp, out, msgs in rounds:
        assert out.numpy().tobytes() == msg.numpy().tobytes()
        e0, e1, e2 = oracle.ac17_cp_encrypt(pk, m, pi, s.numpy().tobytes()[:64], msgs[0])
        assert (c0.numpy().tobytes()[:384], c.numpy().tobytes()[:192 * n1], cp.numpy().tobytes()[:384]) == (e0, e1, e2)
    A.set_async(False); Bc.set_async(False)
