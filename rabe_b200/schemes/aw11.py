"""AW11 (Lewko-Waters decentralised) CP-ABE with rabe's API shape
(/root/reference/src/schemes/aw11/mod.rs) over the GPU C ABI."""
from dataclasses import dataclass
from typing import List, Optional, Tuple

from ..error import RabeError
from ..policy import Policy, PolicyLanguage, remove_index, sha3_hash_fr
from .common import G1_GEN, G2_GEN, TABLES, Rng, chunks, decrypt_symmetric, encrypt_symmetric, engine, u8


@dataclass
class Aw11GlobalKey:            # aw11/mod.rs:50
    g1: bytes
    g2: bytes


@dataclass
class Aw11PublicKey:            # aw11/mod.rs:59
    attr: List[Tuple[str, bytes, bytes]]


@dataclass
class Aw11MasterKey:            # aw11/mod.rs:67
    attr: List[Tuple[str, bytes, bytes]]


@dataclass
class Aw11Ciphertext:           # aw11/mod.rs:75
    policy: Tuple[str, PolicyLanguage]
    c_0: bytes
    c: List[Tuple[str, bytes, bytes, bytes]]
    ct: bytes


@dataclass
class Aw11SecretKey:            # aw11/mod.rs:86
    gid: str
    attr: List[Tuple[str, bytes]]


_EGG = {}


def _e_gg(gk: Aw11GlobalKey) -> bytes:
    key = gk.g1 + gk.g2
    if key not in _EGG:
        _EGG[key] = engine().pairing(u8(gk.g1), u8(gk.g2)).tobytes()
    return _EGG[key]


def setup(rng: Rng = None) -> Aw11GlobalKey:
    """aw11/mod.rs:100-108."""
    rng = rng or Rng()
    e = engine()
    return Aw11GlobalKey(e.g1_mul_var(u8(G1_GEN), u8(rng.fr())).tobytes(), e.g2_mul_var(u8(G2_GEN), u8(rng.fr())).tobytes())


def authgen(gk: Aw11GlobalKey, attributes: List[str], rng: Rng = None) -> Optional[Tuple[Aw11PublicKey, Aw11MasterKey]]:
    """aw11/mod.rs:121-151."""
    if len(attributes) == 0:
        return None
    rng = rng or Rng()
    e = engine()
    n = len(attributes)
    draws = chunks(rng.frs(2 * n), 32)                               # alpha_i, y_i interleaved
    alphas, ys = b"".join(draws[0::2]), b"".join(draws[1::2])
    gts = e.gt_pow_fixed(TABLES.get("gt", _e_gg(gk), 8), u8(alphas)).tobytes()
    g2s = e.g2_mul_fixed(TABLES.get("g2", gk.g2, 8), u8(ys)).tobytes()
    names = [a.upper() for a in attributes]
    pk = Aw11PublicKey([(nm, gts[384 * i:384 * i + 384], g2s[128 * i:128 * i + 128]) for i, nm in enumerate(names)])
    msk = Aw11MasterKey([(nm, draws[2 * i], draws[2 * i + 1]) for i, nm in enumerate(names)])
    return pk, msk


def add_to_attribute(gk: Aw11GlobalKey, msk: Aw11MasterKey, attribute: str, sk: Aw11SecretKey):
    """aw11/mod.rs:200-232:  K = g1*alpha + H(gid)*g1*y = g1*(alpha + H(gid) y)."""
    if len(attribute) == 0:
        raise RabeError("empty _attributes")
    if len(sk.gid) == 0:
        raise RabeError("empty _gid")
    e = engine()
    auth = next((x for x in msk.attr if x[0] == attribute), None)
    if auth is None:
        raise RabeError("attribute not held by this authority")        # the reference unwrap()s (panic)
    sc = e.fr_op("add", u8(auth[1]), e.fr_op("mul", u8(sha3_hash_fr(sk.gid)), u8(auth[2])))
    sk.attr.append((auth[0].upper(), e.g1_mul_fixed(TABLES.get("g1", gk.g1, 16), sc).tobytes()))


def keygen(gk: Aw11GlobalKey, msk: Aw11MasterKey, name: str, attributes: List[str]) -> Aw11SecretKey:
    """aw11/mod.rs:165-190."""
    if len(attributes) == 0:
        raise RabeError("empty _attributes")
    if len(name) == 0:
        raise RabeError("empty _name")
    sk = Aw11SecretKey(name, [])
    for a in attributes:
        add_to_attribute(gk, msk, a, sk)
    return sk


def find_pk_attr(pks: List[Aw11PublicKey], attr: str):           # aw11/mod.rs:374
    for pk in pks:
        for t in pk.attr:
            if t[0] == attr:
                return t
    return None


def encrypt(gk: Aw11GlobalKey, pks: List[Aw11PublicKey], policy: str, language: PolicyLanguage, data: bytes, rng: Rng = None, _msg=None):
    """aw11/mod.rs:241-289."""
    rng = rng or Rng()
    e = engine()
    pol = Policy(policy, language)
    pol.msp()                                                        # :253 (binary ANDs only; error instead of panic)
    plan = e.share_plan(pol)
    labels = pol.leaf_labels()
    n = plan.n_leaves
    s = rng.fr()
    s_coeffs = rng.frs(plan.n_coefs)
    w_coeffs = rng.frs(plan.n_coefs)
    gt_tab = TABLES.get("gt", _e_gg(gk), 8)
    msg = _msg if _msg is not None else e.gt_pow_fixed(gt_tab, u8(rng.fr())).tobytes()
    r_x = chunks(rng.frs(n), 32)                                     # drawn for every share, found or not
    s_shares = chunks(e.shares(plan, u8(s), u8(s_coeffs)), 32)
    w_shares = chunks(e.shares(plan, u8(b"\0" * 32), u8(w_coeffs)), 32)
    c_0 = e.gt_mul(u8(msg), e.gt_pow_fixed(gt_tab, u8(s))).tobytes()
    rows = [(i, l, find_pk_attr(pks, remove_index(l.upper()))) for i, l in enumerate(labels)]
    rows = [(i, l, a) for i, l, a in rows if a is not None]
    c = []
    if rows and len(rows) == n:
        # every leaf has an authority key: one fused call (rb_aw11_encrypt_batch) recomputes the shares on the device
        pk_gt, pk_g2 = u8(b"".join(a[1] for _, _, a in rows)), u8(b"".join(a[2] for _, _, a in rows))
        c0f, c1, c2, c3 = [x.tobytes() for x in e.aw11_encrypt(TABLES.get("g2", gk.g2, 8), gt_tab, plan, pk_gt, pk_g2, u8(s), u8(s_coeffs),
                                                                u8(w_coeffs), u8(b"".join(r_x)), u8(msg))]
        assert c0f == c_0
        c = [(l.upper(), c1[384 * k:384 * k + 384], c2[128 * k:128 * k + 128], c3[128 * k:128 * k + 128]) for k, (_, l, _) in enumerate(rows)]
    elif rows:
        sh = u8(b"".join(s_shares[i] for i, _, _ in rows))
        rx = u8(b"".join(r_x[i] for i, _, _ in rows))
        ws = u8(b"".join(w_shares[i] for i, _, _ in rows))
        c1 = e.gt_mul(e.gt_pow_fixed(gt_tab, sh), e.gt_pow_var(u8(b"".join(a[1] for _, _, a in rows)), rx)).tobytes()
        g2t = TABLES.get("g2", gk.g2, 8)
        c2 = e.g2_mul_fixed(g2t, rx).tobytes()
        c3 = e.g2_add(e.g2_mul_var(u8(b"".join(a[2] for _, _, a in rows)), rx), e.g2_mul_fixed(g2t, ws)).tobytes()
        c = [(l.upper(), c1[384 * k:384 * k + 384], c2[128 * k:128 * k + 128], c3[128 * k:128 * k + 128]) for k, (_, l, _) in enumerate(rows)]
    return Aw11Ciphertext((policy, PolicyLanguage(language)), c_0, c, encrypt_symmetric(msg, data, rng))


def decrypt_gt(gk: Aw11GlobalKey, sk: Aw11SecretKey, ct: Aw11Ciphertext) -> bytes:
    e = engine()
    str_attr = [x[0] for x in sk.attr]
    pol = Policy(ct.policy[0], ct.policy[1])
    if not pol.satisfied(str_attr):
        raise RabeError("Error: attributes in sk do not match policy in ct.")
    ok, pruned = pol.prune(str_attr)
    if not ok:
        raise RabeError("Error in aw11/decrypt: attributes in sk do not match policy in ct.")
    labels = pol.leaf_labels()
    coeffs = chunks(e.policy_coefficients(pol, len(labels)), 32)
    h = e.g1_mul_fixed(TABLES.get("g1", gk.g1, 16), u8(sha3_hash_fr(sk.gid))).tobytes()       # sha3_hash(g1, gid)
    sk_names, ct_names = [x[0] for x in sk.attr], [x[0] for x in ct.c]
    ct_idx, sk_idx, coeff = [], [], b""
    for name, label in pruned:
        sk_idx.append(sk_names.index(name)); ct_idx.append(ct_names.index(label))
        coeff += next(cv for l, cv in zip(labels, coeffs) if l == label)
    # rb_aw11_decrypt_batch: msg = c_0 * prod C1^-c * e(-c h, C3) * e(c K, C2), one final exponentiation
    return e.aw11_decrypt(u8(h), u8(b"".join(x[1] for x in sk.attr)), u8(ct.c_0), u8(b"".join(x[1] for x in ct.c)),
                          u8(b"".join(x[2] for x in ct.c)), u8(b"".join(x[3] for x in ct.c)), ct_idx, sk_idx, u8(coeff)).tobytes()


def decrypt(gk: Aw11GlobalKey, sk: Aw11SecretKey, ct: Aw11Ciphertext) -> bytes:
    """aw11/mod.rs:298-372."""
    return decrypt_symmetric(decrypt_gt(gk, sk, ct), ct.ct)
