"""BSW CP-ABE with rabe's API shape (/root/reference/src/schemes/bsw/mod.rs); every field/group
operation runs on the GPU through the C ABI (fixed-base tables for pk members, batched variable-base
multiplications, one pairing product with a single final exponentiation per decryption)."""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from ..error import RabeError
from ..policy import Policy, PolicyLanguage, remove_index, sha3_hash_fr
from .common import (FR_MINUS_ONE, G1_GEN, G2_GEN, HANDLES, TABLES, Rng, chunks, decrypt_symmetric, encrypt_symmetric, engine, u8)


@dataclass
class CpAbePublicKey:           # bsw/mod.rs:43
    g1: bytes
    g2: bytes
    h: bytes
    f: bytes
    e_gg_alpha: bytes


@dataclass
class CpAbeMasterKey:           # bsw/mod.rs:55
    beta: bytes
    g2_alpha: bytes


@dataclass
class CpAbeAttribute:           # bsw/mod.rs:85
    string: str
    g1: bytes
    g2: bytes


@dataclass
class CpAbeCiphertext:          # bsw/mod.rs:64
    policy: Tuple[str, PolicyLanguage]
    c: bytes
    c_p: bytes
    c_y: List[CpAbeAttribute]
    data: bytes


@dataclass
class CpAbeSecretKey:           # bsw/mod.rs:76
    d: bytes
    d_j: List[CpAbeAttribute]


def _pk_handle(pk: "CpAbePublicKey"):
    """Device tables of the public key for the fused entry points (built once per key and engine)."""
    return HANDLES.get("bsw_pk", pk.g1 + pk.g2 + pk.h + pk.e_gg_alpha, lambda e: e.bsw_pk_load(pk.g1, pk.g2, pk.h, pk.e_gg_alpha))


def setup(rng: Rng = None) -> Tuple[CpAbePublicKey, CpAbeMasterKey]:
    """bsw/mod.rs:92-115."""
    rng = rng or Rng()
    e = engine()
    g1 = e.g1_mul_var(u8(G1_GEN), u8(rng.fr())).tobytes()
    g2 = e.g2_mul_var(u8(G2_GEN), u8(rng.fr())).tobytes()
    beta, alpha = rng.fr(), rng.fr()
    h = e.g1_mul_var(u8(g1), u8(beta)).tobytes()
    f = e.g2_mul_var(u8(g2), e.fr_op("inverse", u8(beta))).tobytes()
    g2_alpha = e.g2_mul_var(u8(g2), u8(alpha)).tobytes()
    e_gg_alpha = e.pairing(u8(g1), u8(g2_alpha)).tobytes()
    return CpAbePublicKey(g1, g2, h, f, e_gg_alpha), CpAbeMasterKey(beta, g2_alpha)


def keygen(pk: CpAbePublicKey, msk: CpAbeMasterKey, attributes: List[str], rng: Rng = None) -> Optional[CpAbeSecretKey]:
    """bsw/mod.rs:125-152."""
    if len(attributes) == 0:
        return None
    rng = rng or Rng()
    e = engine()
    n = len(attributes)
    r = rng.fr()
    r_j = rng.frs(n)
    hashes = b"".join(sha3_hash_fr(j) for j in attributes)
    d, dj_g1, dj_g2 = [x.tobytes() for x in e.bsw_keygen(_pk_handle(pk), u8(msk.beta), u8(msk.g2_alpha), u8(hashes), u8(r), u8(r_j))]   # rb_bsw_keygen_batch
    return CpAbeSecretKey(d, [CpAbeAttribute(a, dj_g1[64 * i:64 * i + 64], dj_g2[128 * i:128 * i + 128]) for i, a in enumerate(attributes)])


def delegate(pk: CpAbePublicKey, sk: CpAbeSecretKey, subset: List[str], rng: Rng = None) -> Optional[CpAbeSecretKey]:
    """bsw/mod.rs:162-206."""
    names = [x.string for x in sk.d_j]
    if not set(subset) <= set(names) or len(subset) == 0:
        return None
    rng = rng or Rng()
    e = engine()
    r = rng.fr()
    r_j = rng.frs(len(subset))
    src = [next(x for x in sk.d_j if x.string == a) for a in subset]
    hashes = b"".join(sha3_hash_fr(a) for a in subset)
    d, g1n, g2n = [x.tobytes() for x in e.bsw_delegate(_pk_handle(pk), u8(pk.f), u8(sk.d), u8(b"".join(x.g1 for x in src)), u8(b"".join(x.g2 for x in src)),
                                                       u8(hashes), u8(r), u8(r_j))]                                       # rb_bsw_delegate_batch
    return CpAbeSecretKey(d, [CpAbeAttribute(a, g1n[64 * i:64 * i + 64], g2n[128 * i:128 * i + 128]) for i, a in enumerate(subset)])


def encrypt(pk: CpAbePublicKey, policy: str, language: PolicyLanguage, plaintext: bytes, rng: Rng = None, _msg=None) -> CpAbeCiphertext:
    """bsw/mod.rs:217-251."""
    return encrypt_batch(pk, policy, language, [plaintext], rng, _msgs=None if _msg is None else [_msg])[0]


def encrypt_batch(pk, policy, language, plaintexts, rng: Rng = None, _msgs=None) -> List[CpAbeCiphertext]:
    rng = rng or Rng()
    e = engine()
    pol = Policy(policy, language)
    plan = e.share_plan(pol)
    labels = pol.leaf_labels()
    B, n = len(plaintexts), plan.n_leaves
    secrets, coeffs, rho = b"", b"", b""
    for _ in range(B):                       # draw order of the reference: secret, msg, share coefficients
        secrets += rng.fr()
        if _msgs is None:
            rho += rng.fr()
        coeffs += rng.frs(plan.n_coefs)
    gt_tab = TABLES.get("gt", pk.e_gg_alpha, 8)
    msgs = e.gt_pow_fixed(gt_tab, u8(rho)).tobytes() if _msgs is None else b"".join(_msgs)
    leaf_hash = b"".join(sha3_hash_fr(remove_index(l)) for l in labels)
    c, c_p, cy_g1, cy_g2 = [x.tobytes() for x in e.bsw_encrypt(_pk_handle(pk), plan, u8(leaf_hash), u8(secrets), u8(coeffs), u8(msgs))]   # rb_bsw_encrypt_batch
    out = []
    for b in range(B):
        c_y = [CpAbeAttribute(l, cy_g1[64 * (b * n + i):64 * (b * n + i + 1)], cy_g2[128 * (b * n + i):128 * (b * n + i + 1)]) for i, l in enumerate(labels)]
        out.append(CpAbeCiphertext((policy, PolicyLanguage(language)), c[64 * b:64 * b + 64], c_p[384 * b:384 * b + 384], c_y,
                                   encrypt_symmetric(msgs[384 * b:384 * b + 384], plaintexts[b], rng)))
    return out


def decrypt_gt(sk: CpAbeSecretKey, ct: CpAbeCiphertext) -> bytes:
    """The Gt value `_msg` of bsw/mod.rs:308 (before the KEM)."""
    e = engine()
    attr = [x.string for x in sk.d_j]
    pol = Policy(ct.policy[0], ct.policy[1])
    if not pol.satisfied(attr):
        raise RabeError("Error in bsw/encrypt: attributes do not match policy.")
    ok, pruned = pol.prune(attr)
    if not ok:
        raise RabeError("Error in bsw/encrypt: attributes do not match policy.")
    labels = pol.leaf_labels()
    z = chunks(e.policy_coefficients(pol, len(labels)), 32)                       # calc_coefficients
    ct_names, sk_names = [x.string for x in ct.c_y], [x.string for x in sk.d_j]
    ct_idx, sk_idx, coeff = [], [], b""
    for k, j in pruned:                      # bsw/mod.rs:282-307: leaves missing on either side are skipped
        if j not in ct_names or k not in sk_names:
            continue
        for label, zc in zip(labels, z):
            if label == j:
                ct_idx.append(ct_names.index(j)); sk_idx.append(sk_names.index(k)); coeff += zc
    # rb_bsw_decrypt_batch: prod e(z c_y.g1, d_j.g2) e(-z d_j.g1, c_y.g2) * e(-c, d), one final exponentiation, times c_p
    return e.bsw_decrypt(u8(sk.d), u8(b"".join(x.g1 for x in sk.d_j)), u8(b"".join(x.g2 for x in sk.d_j)), u8(ct.c), u8(ct.c_p),
                         u8(b"".join(x.g1 for x in ct.c_y)), u8(b"".join(x.g2 for x in ct.c_y)), ct_idx, sk_idx, u8(coeff)).tobytes()


def decrypt(sk: CpAbeSecretKey, ct: CpAbeCiphertext) -> bytes:
    """bsw/mod.rs:260-318."""
    return decrypt_symmetric(decrypt_gt(sk, ct), ct.data)
