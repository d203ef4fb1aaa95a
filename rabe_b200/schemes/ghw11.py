"""GHW11 CP-ABE with outsourced decryption, rabe's API shape (/root/reference/src/schemes/ghw11/mod.rs)
over the GPU C ABI.  The batch-server half -- `transform` -- is one fused call (rb_ghw11_transform_batch:
every pairing has a key-side G2 argument, so all 2 nI + 1 Miller loops of an item run from line tables
built once per call and share one final exponentiation); key generation and encryption compose the
batched fixed-base operators like the other scheme mirrors."""
from dataclasses import dataclass
from typing import List, Optional, Tuple

from ..error import RabeError
from ..policy import Policy, PolicyLanguage, remove_index, sha3_hash_fr
from .common import G1_GEN, G2_GEN, TABLES, Rng, chunks, decrypt_symmetric, encrypt_symmetric, engine, u8


@dataclass
class Ghw11PublicKey:           # ghw11/mod.rs:22
    g1: bytes
    g2: bytes
    g1_a: bytes
    g2_a: bytes
    e_gg_alpha: bytes


@dataclass
class Ghw11MasterKey:           # ghw11/mod.rs:33
    g2_alpha: bytes
    pk: Ghw11PublicKey


@dataclass
class Ghw11Attribute:           # ghw11/mod.rs:53
    string: str
    k_x: bytes


@dataclass
class Ghw11SecretKey:           # ghw11/mod.rs:41
    k: bytes
    l: bytes
    attr_key: List[Ghw11Attribute]


@dataclass
class Ghw11TransformKey:        # ghw11/mod.rs:61
    k_z: bytes
    l_z: bytes
    attr_key_z: List[Ghw11Attribute]


@dataclass
class Ghw11RetrieveKey:         # ghw11/mod.rs:70
    z: bytes


@dataclass
class Ghw11Ciphertext:          # ghw11/mod.rs:77
    policy: Tuple[str, PolicyLanguage]
    c: bytes
    c1: bytes
    ci_di: List[Tuple[str, bytes, bytes]]
    data: bytes


@dataclass
class Ghw11TransformCiphertext:  # ghw11/mod.rs:88
    c: bytes
    t: bytes


def setup(rng: Rng = None) -> Tuple[Ghw11PublicKey, Ghw11MasterKey]:
    """ghw11/mod.rs:92-111."""
    rng = rng or Rng()
    e = engine()
    g1 = e.g1_mul_var(u8(G1_GEN), u8(rng.fr())).tobytes()
    g2 = e.g2_mul_var(u8(G2_GEN), u8(rng.fr())).tobytes()
    a = rng.fr()
    g1_a = e.g1_mul_var(u8(g1), u8(a)).tobytes()
    g2_a = e.g2_mul_var(u8(g2), u8(a)).tobytes()
    alpha = rng.fr()
    e_gg_alpha = e.gt_pow_var(e.pairing(u8(g1), u8(g2)), u8(alpha)).tobytes()
    g2_alpha = e.g2_mul_var(u8(g2), u8(alpha)).tobytes()
    pk = Ghw11PublicKey(g1, g2, g1_a, g2_a, e_gg_alpha)
    return pk, Ghw11MasterKey(g2_alpha, pk)


def keygen(pk: Ghw11PublicKey, msk: Ghw11MasterKey, attributes: List[str], rng: Rng = None) -> Optional[Ghw11SecretKey]:
    """ghw11/mod.rs:121-151."""
    if len(attributes) == 0:
        return None
    rng = rng or Rng()
    e = engine()
    r = rng.fr()
    g2t = TABLES.get("g2", pk.g2, 8)
    l = e.g2_mul_fixed(g2t, u8(r)).tobytes()
    k = e.g2_add(e.g2_mul_fixed(TABLES.get("g2", pk.g2_a, 8), u8(r)), u8(msk.g2_alpha)).tobytes()
    hashes = u8(b"".join(sha3_hash_fr(j) for j in attributes))
    kx = e.g2_mul_fixed(g2t, e.fr_op("mul", hashes, u8(r))).tobytes()                # sha3_hash(g2, j) * r = g2 * (H(j) r)
    return Ghw11SecretKey(k, l, [Ghw11Attribute(j, kx[128 * i:128 * i + 128]) for i, j in enumerate(attributes)])


def tkgen(sk: Ghw11SecretKey, rng: Rng = None) -> Optional[Tuple[Ghw11TransformKey, Ghw11RetrieveKey]]:
    """ghw11/mod.rs:156-179."""
    rng = rng or Rng()
    e = engine()
    z = rng.fr()
    zi = e.fr_op("inverse", u8(z))
    pts = sk.k + sk.l + b"".join(x.k_x for x in sk.attr_key)
    out = e.g2_mul_var(u8(pts), u8(zi.tobytes() * (len(pts) // 128))).tobytes()
    attr = [Ghw11Attribute(x.string, out[128 * (2 + i):128 * (3 + i)]) for i, x in enumerate(sk.attr_key)]
    return Ghw11TransformKey(out[:128], out[128:256], attr), Ghw11RetrieveKey(z)


def encrypt(pk: Ghw11PublicKey, policy: str, language: PolicyLanguage, plaintext: bytes, rng: Rng = None, _msg=None) -> Ghw11Ciphertext:
    """ghw11/mod.rs:190-224.  Draw order of the reference: secret, msg, share coefficients, t_i per share."""
    rng = rng or Rng()
    e = engine()
    pol = Policy(policy, language)
    plan = e.share_plan(pol)
    labels = pol.leaf_labels()
    secret = rng.fr()
    gt_tab = TABLES.get("gt", pk.e_gg_alpha, 8)
    msg = _msg if _msg is not None else e.gt_pow_fixed(gt_tab, u8(rng.fr())).tobytes()
    coeffs = rng.frs(plan.n_coefs)
    t_i = rng.frs(plan.n_leaves)
    shares = e.shares(plan, u8(secret), u8(coeffs))
    g1t = TABLES.get("g1", pk.g1, 16)
    c = e.gt_mul(e.gt_pow_fixed(gt_tab, u8(secret)), u8(msg)).tobytes()
    c1 = e.g1_mul_fixed(g1t, u8(secret)).tobytes()
    hashes = u8(b"".join(sha3_hash_fr(remove_index(l)) for l in labels))
    # C_i = g1_a * share + H(attr) g1 * (-t_i) ; D_i = g1 * t_i
    ci = e.g1_add(e.g1_mul_fixed(TABLES.get("g1", pk.g1_a, 16), shares),
                  e.g1_mul_fixed(g1t, e.fr_op("neg", e.fr_op("mul", hashes, u8(t_i))))).tobytes()
    di = e.g1_mul_fixed(g1t, u8(t_i)).tobytes()
    ci_di = [(l, ci[64 * i:64 * i + 64], di[64 * i:64 * i + 64]) for i, l in enumerate(labels)]
    return Ghw11Ciphertext((policy, PolicyLanguage(language)), c, c1, ci_di, encrypt_symmetric(msg, plaintext, rng))


def transform_batch(cts: List[Ghw11Ciphertext], tk: Ghw11TransformKey) -> List[Ghw11TransformCiphertext]:
    """ghw11::transform for B ciphertexts of ONE policy under one transform key: a single rb_ghw11_transform_batch."""
    e = engine()
    attr = [x.string for x in tk.attr_key_z]
    pol = Policy(cts[0].policy[0], cts[0].policy[1])
    if any(ct.policy != cts[0].policy for ct in cts):
        raise RabeError("transform_batch: the ciphertexts of a batch share one policy")
    if not pol.satisfied(attr):
        raise RabeError("Error: attributes in tk do not match policy in ct.")
    ok, pruned = pol.prune(attr)
    if not ok:
        raise RabeError("Error in Ghw11/decrypt: attributes in sk do not match policy in ct.")
    labels = pol.leaf_labels()
    coeffs = chunks(e.policy_coefficients(pol, len(labels)), 32)
    ct_names = [x[0] for x in cts[0].ci_di]
    ct_idx, sk_idx, coeff = [], [], b""
    for name, label in pruned:
        sk_idx.append(attr.index(name)); ct_idx.append(ct_names.index(label))
        coeff += next(cv for l, cv in zip(labels, coeffs) if l == label)
    t = e.ghw11_transform(u8(tk.k_z), u8(tk.l_z), u8(b"".join(x.k_x for x in tk.attr_key_z)), u8(b"".join(ct.c1 for ct in cts)),
                          u8(b"".join(x[1] for ct in cts for x in ct.ci_di)), u8(b"".join(x[2] for ct in cts for x in ct.ci_di)),
                          ct_idx, sk_idx, u8(coeff)).tobytes()
    return [Ghw11TransformCiphertext(ct.c, t[384 * b:384 * b + 384]) for b, ct in enumerate(cts)]


def transform(ct: Ghw11Ciphertext, tk: Ghw11TransformKey) -> Ghw11TransformCiphertext:
    """ghw11/mod.rs:227-294."""
    return transform_batch([ct], tk)[0]


def decrypt_out_gt(pct: Ghw11TransformCiphertext, rk: Ghw11RetrieveKey) -> bytes:
    return engine().ghw11_decrypt_out(u8(pct.c), u8(pct.t), u8(rk.z)).tobytes()


def decrypt_out(pct: Ghw11TransformCiphertext, rk: Ghw11RetrieveKey, data: bytes) -> bytes:
    """ghw11/mod.rs:297-305."""
    return decrypt_symmetric(decrypt_out_gt(pct, rk), data)
