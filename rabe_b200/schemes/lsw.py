"""LSW KP-ABE with rabe's API shape (/root/reference/src/schemes/lsw/mod.rs) over the GPU C ABI."""
from dataclasses import dataclass
from typing import List, Tuple

from ..error import RabeError
from ..policy import Policy, PolicyLanguage, remove_index, sha3_hash_fr
from .common import G1_GEN, G2_GEN, HANDLES, TABLES, Rng, chunks, decrypt_symmetric, encrypt_symmetric, engine, u8

G1_ZERO, G2_ZERO = b"\0" * 64, b"\0" * 128


@dataclass
class KpAbePublicKey:           # lsw/mod.rs:44
    g1: bytes
    g2: bytes
    g1_b: bytes
    g1_b2: bytes
    h_b: bytes
    e_gg_alpha: bytes


@dataclass
class KpAbeMasterKey:           # lsw/mod.rs:57
    alpha1: bytes
    alpha2: bytes
    b: bytes
    h_g1: bytes
    h_g2: bytes


@dataclass
class KpAbeSecretKey:           # lsw/mod.rs:69
    policy: Tuple[str, PolicyLanguage]
    dj: List[Tuple[str, bytes, bytes, bytes, bytes, bytes]]


@dataclass
class KpAbeCiphertext:          # lsw/mod.rs:78
    e1: bytes
    e2: bytes
    ej: List[Tuple[str, bytes, bytes, bytes]]
    ct: bytes


def is_negative(attr: str) -> bool:     # tools/mod.rs:6
    return attr[:1] == "!"


def setup(rng: Rng = None):
    """lsw/mod.rs:86-110."""
    rng = rng or Rng()
    e = engine()
    alpha1, alpha2, b = rng.fr(), rng.fr(), rng.fr()
    g1 = e.g1_mul_var(u8(G1_GEN), u8(rng.fr())).tobytes()
    g2 = e.g2_mul_var(u8(G2_GEN), u8(rng.fr())).tobytes()
    h_g1 = e.g1_mul_var(u8(G1_GEN), u8(rng.fr())).tobytes()
    h_g2 = e.g2_mul_var(u8(G2_GEN), u8(rng.fr())).tobytes()
    g1_b = e.g1_mul_var(u8(g1), u8(b)).tobytes()
    g1_b2 = e.g1_mul_var(u8(g1_b), u8(b)).tobytes()
    h_b = e.g1_mul_var(u8(h_g1), u8(b)).tobytes()
    e_gg_alpha = e.gt_pow_var(e.pairing(u8(g1), u8(g2)), e.fr_op("mul", u8(alpha1), u8(alpha2))).tobytes()
    return KpAbePublicKey(g1, g2, g1_b, g1_b2, h_b, e_gg_alpha), KpAbeMasterKey(alpha1, alpha2, b, h_g1, h_g2)


def keygen(pk: KpAbePublicKey, msk: KpAbeMasterKey, policy: str, language: PolicyLanguage, rng: Rng = None) -> KpAbeSecretKey:
    """lsw/mod.rs:121-170."""
    rng = rng or Rng()
    e = engine()
    pol = Policy(policy, language)
    plan = e.share_plan(pol)
    labels = pol.leaf_labels()
    coeffs = rng.frs(plan.n_coefs)                                  # gen_shares draws first ...
    rand = rng.frs(plan.n_leaves)                                   # ... then `random` per leaf
    names = [remove_index(l) for l in labels]
    hashes = u8(b"".join(sha3_hash_fr(n) for n in names))
    g1t, g2t = TABLES.get("g1", pk.g1, 16), TABLES.get("g2", pk.g2, 8)
    # positive leaves: (g1*(alpha2*share) + H(attr)*g1*random, g2*random)          rb_lsw_keygen_batch
    d1, d2 = [x.tobytes() for x in e.lsw_keygen(g1t, g2t, plan, hashes, u8(msk.alpha1), u8(msk.alpha2), u8(coeffs), u8(rand))]
    shares = e.shares(plan, u8(msk.alpha1), u8(coeffs)) if any(is_negative(n) for n in names) else None
    dj = []
    for i, n in enumerate(names):
        if is_negative(n):
            sh, rd, hh = shares.tobytes()[32 * i:32 * i + 32], rand[32 * i:32 * i + 32], hashes.tobytes()[32 * i:32 * i + 32]
            d3 = e.g1_add(e.g1_mul_fixed(g1t, u8(sh)), e.g1_mul_var(u8(pk.g1_b2), u8(rd))).tobytes()
            d4 = e.g1_add(e.g1_mul_var(u8(pk.g1_b), e.fr_op("mul", u8(hh), u8(rd))), e.g1_mul_var(u8(msk.h_g1), u8(rd))).tobytes()
            d5 = e.g1_mul_fixed(g1t, e.fr_op("neg", u8(rd))).tobytes()
            dj.append((n, G1_ZERO, G2_ZERO, d3, d4, d5))
        else:
            dj.append((n, d1[64 * i:64 * i + 64], d2[128 * i:128 * i + 128], G1_ZERO, G1_ZERO, G1_ZERO))
    return KpAbeSecretKey((policy, PolicyLanguage(language)), dj)


def _pk_handle(pk: KpAbePublicKey):
    """Fixed-base tables of every public-key member lsw::encrypt multiplies (rb_lsw_pk_load), once per key and engine."""
    return HANDLES.get("lsw_pk", pk.g1 + pk.g2 + pk.g1_b + pk.g1_b2 + pk.h_b + pk.e_gg_alpha,
                       lambda e: e.lsw_pk_load(pk.g1, pk.g2, pk.g1_b, pk.g1_b2, pk.h_b, pk.e_gg_alpha))


def encrypt_batch(pk: KpAbePublicKey, attributes: List[str], plaintexts: List[bytes], rng: Rng = None, _msgs=None) -> List[KpAbeCiphertext]:
    """B independent lsw::encrypt calls over one attribute list in ONE fused call (rb_lsw_encrypt_batch).  Randomness
    is drawn per item in the reference's order: secret, the n `sx` draws, (msg) -- lsw/mod.rs:193-210."""
    if len(attributes) == 0 or any(len(p) == 0 for p in plaintexts):
        raise RabeError("attributes or data empty")
    rng = rng or Rng()
    e = engine()
    n, B = len(attributes), len(plaintexts)
    gt_tab = TABLES.get("gt", pk.e_gg_alpha, 8)
    secrets, draws, msgs = b"", b"", []
    for b in range(B):
        secrets += rng.fr()
        draws += rng.frs(n)                                          # pushed as sx[1..n]; the `sx[0]` quirk is reproduced on the device
        msgs.append(_msgs[b] if _msgs is not None else e.gt_pow_fixed(gt_tab, u8(rng.fr())).tobytes())
    hashes = u8(b"".join(sha3_hash_fr(a) for a in attributes))
    e1, e2, j1, j2, j3 = [x.tobytes() for x in e.lsw_encrypt(_pk_handle(pk), hashes, u8(secrets), u8(draws), u8(b"".join(msgs)))]
    out = []
    for b in range(B):
        ej = [(a, j1[64 * (b * n + i):64 * (b * n + i + 1)], j2[64 * (b * n + i):64 * (b * n + i + 1)], j3[64 * (b * n + i):64 * (b * n + i + 1)])
              for i, a in enumerate(attributes)]
        out.append(KpAbeCiphertext(e1[384 * b:384 * b + 384], e2[128 * b:128 * b + 128], ej, encrypt_symmetric(msgs[b], plaintexts[b], rng)))
    return out


def encrypt(pk: KpAbePublicKey, attributes: List[str], plaintext: bytes, rng: Rng = None, _msg=None) -> KpAbeCiphertext:
    """lsw/mod.rs:180-219 (including the `sx[0]` quirk at :197-200)."""
    if len(attributes) == 0 or len(plaintext) == 0:
        raise RabeError("attributes or data empty")
    return encrypt_batch(pk, attributes, [plaintext], rng, None if _msg is None else [_msg])[0]


def decrypt_gt(sk: KpAbeSecretKey, ct: KpAbeCiphertext) -> bytes:
    e = engine()
    attr = [x[0] for x in ct.ej]
    pol = Policy(sk.policy[0], sk.policy[1])
    ok, pruned = pol.prune(attr)
    if not ok:
        raise RabeError("Error in lsw/decrypt: attributes do not match policy.")
    labels = pol.leaf_labels()
    coeffs = chunks(e.policy_coefficients(pol, len(labels)), 32)
    sk_names, ct_names = [x[0] for x in sk.dj], [x[0] for x in ct.ej]
    ct_idx, sk_idx, coeff = [], [], b""
    for name, label in pruned:
        if is_negative(name):
            raise RabeError("lsw/decrypt: negative attributes are not decryptable (TODO in the reference, lsw/mod.rs:265-273)")
        sk_idx.append(sk_names.index(name)); ct_idx.append(ct_names.index(name))
        coeff += next(cv for l, cv in zip(labels, coeffs) if l == label)
    # rb_lsw_decrypt_batch: msg = e1 * prod e(-c d1, e2) * e(c E1, d2), one final exponentiation
    return e.lsw_decrypt(u8(b"".join(x[1] for x in sk.dj)), u8(b"".join(x[2] for x in sk.dj)), u8(ct.e1), u8(ct.e2),
                         u8(b"".join(x[1] for x in ct.ej)), ct_idx, sk_idx, u8(coeff)).tobytes()


def decrypt(sk: KpAbeSecretKey, ct: KpAbeCiphertext) -> bytes:
    """lsw/mod.rs:228-290."""
    return decrypt_symmetric(decrypt_gt(sk, ct), ct.ct)
