"""The restructured AC17 algorithm on the CPU (oracle/ac17_fast.cpp: fixed-base window tables, folded policy scalars,
one final exponentiation) reproduces the reference-sequence restatement (oracle/ac17.cpp) byte for byte -- the CPU twin
of the claim the CUDA path makes, and the second CPU mode bench.py times."""
import random

import oracle
from oracle import policy as opol
import rb_testutil as util
from rb_testutil import rand_fr


def test_fast_mode_equals_reference_sequence():
    rng = random.Random(51)
    pk, msk = oracle.ac17_setup(rand_fr(rng, 9))
    for policy, attrs in (('("A" and "B") and ("C" and "D")', ["A", "B", "C", "D"]),
                          ('("A" or "X") and ("C" or ("Y" and "B"))', ["Q", "C", "A", "B"]),
                          ('"A"', ["A"])):
        tree = opol.parse(policy, opol.HUMAN)
        m, pi, n2 = opol.calculate_msp(tree)
        fast = oracle.Ac17Fast(pk, m, pi)
        k0, k, kp = oracle.ac17_cp_keygen(msk, attrs, rand_fr(rng, len(attrs) + 3))
        ok, pruned = opol.calc_pruned(attrs, tree)
        ct_idx, sk_idx = util.decrypt_lists(pruned, pi, attrs)
        for _ in range(2):
            s, msg = rand_fr(rng, 2), util.gt_random(rng)
            ref = oracle.ac17_cp_encrypt(pk, m, pi, s, msg)
            assert fast.encrypt(s, msg) == ref
            dec = oracle.ac17_cp_decrypt([a for a, _ in pruned], pi, *ref, attrs, k0, k, kp)
            assert oracle.Ac17Fast.decrypt(ct_idx, sk_idx, *ref, k0, k, kp) == dec == msg
        fast.close()
