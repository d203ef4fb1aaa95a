"""Shared helpers for the tests: deterministic synthetic inputs and glue between
the oracle's policy layer and the numeric C ABI."""
import hashlib
import random

import numpy as np

import oracle
from oracle import policy as opol
from oracle import pyref

R = pyref.R
P = pyref.P


def fr(x):
    return int(x % R).to_bytes(32, "big")


def rand_fr(rng, n=1):
    return b"".join(fr(rng.randrange(R)) for _ in range(n))


def sha3_fr(s: str) -> bytes:
    return fr(int.from_bytes(hashlib.sha3_256(s.encode()).digest(), "big"))


def u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


def ac17_hashes(pi, n2):
    """h_row[n1][3][2], h_col[n2][3][2] as the encrypt loops hash them (ac17/mod.rs:305-339)."""
    h_row = b"".join(sha3_fr(f"{name}{l}{t}") for name in pi for l in range(3) for t in range(2))
    h_col = b"".join(sha3_fr(f"0{j + 1}{l}{t}") for j in range(n2) for l in range(3) for t in range(2))
    return h_row, h_col


def ac17_attr_hashes(attrs):
    h_attr = b"".join(sha3_fr(f"{a}{l}{t}") for a in attrs for l in range(3) for t in range(2))
    h_01 = b"".join(sha3_fr(f"01{l}{t}") for l in range(3) for t in range(2))
    return h_attr, h_01


def and_policy(names):
    """left-deep binary AND chain in the human language: ((a0 and a1) and a2) ..."""
    s = f'"{names[0]}"'
    for n in names[1:]:
        s = f'({s} and "{n}")'
    return s


def random_binary_policy(names, rng, p_and=0.5):
    if len(names) == 1:
        return f'"{names[0]}"'
    k = rng.randrange(1, len(names))
    op = "and" if rng.random() < p_and else "or"
    return f"({random_binary_policy(names[:k], rng, p_and)} {op} {random_binary_policy(names[k:], rng, p_and)})"


def gt_random(rng):
    e = oracle.pairing(oracle.g1_generator(), oracle.g2_generator())
    return oracle.gt_pow(e, fr(rng.randrange(R)))


def decrypt_lists(pruned, ct_names, sk_names):
    """Index lists selected by the name-matching loops of ac17/mod.rs:404-413 (every match counts)."""
    ct_idx = [i for cur, _ in pruned for i, n in enumerate(ct_names) if n == cur]
    sk_idx = [i for cur, _ in pruned for i, n in enumerate(sk_names) if n == cur]
    return ct_idx, sk_idx
