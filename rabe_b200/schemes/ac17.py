"""AC17 (FAME) CP-ABE with rabe's API shape (/root/reference/src/schemes/ac17/mod.rs): the same
function names, argument meaning, struct fields (in the same order) and error behaviour; every
group operation runs on the GPU through librabe_b200.so.

Differences that are deliberate and visible:
  * rabe draws randomness from rand::thread_rng() inside each call; here every function takes an
    optional `rng` (see `Rng`) so that runs are reproducible (default: os.urandom-backed).
  * batch forms (`cp_encrypt_batch`, `cp_decrypt_batch`) expose the data-parallel path; the
    single-item functions are batches of one.
Group elements are canonical byte strings (include/rabe_b200.h).
"""
import ctypes
import hashlib
import os
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from .. import _lib
from ..engine import Engine, FR, G1, G2, GT
from ..error import RabeError
from ..policy import Policy, PolicyLanguage, _cstrs

from .common import HANDLES, Rng, engine, set_engine, encrypt_symmetric, decrypt_symmetric, R_ORDER  # noqa: F401


@dataclass
class Ac17PublicKey:            # ac17/mod.rs:62
    g: bytes
    h_a: List[bytes]
    e_gh_ka: List[bytes]

    def to_bytes(self):
        return self.g + b"".join(self.h_a) + b"".join(self.e_gh_ka)

    @staticmethod
    def from_bytes(b):
        return Ac17PublicKey(b[:64], [b[64 + 128 * i:192 + 128 * i] for i in range(3)], [b[448 + 384 * i:832 + 384 * i] for i in range(2)])


@dataclass
class Ac17MasterKey:            # ac17/mod.rs:72
    g: bytes
    h: bytes
    g_k: List[bytes]
    a: List[bytes]
    b: List[bytes]

    def to_bytes(self):
        return self.g + self.h + b"".join(self.g_k) + b"".join(self.a) + b"".join(self.b)

    @staticmethod
    def from_bytes(b):
        return Ac17MasterKey(b[:64], b[64:192], [b[192 + 64 * i:256 + 64 * i] for i in range(3)],
                             [b[384 + 32 * i:416 + 32 * i] for i in range(2)], [b[448 + 32 * i:480 + 32 * i] for i in range(2)])


@dataclass
class Ac17Ciphertext:           # ac17/mod.rs:84
    c_0: List[bytes]
    c: List[Tuple[str, List[bytes]]]
    c_p: bytes
    ct: bytes


@dataclass
class Ac17CpCiphertext:         # ac17/mod.rs:95
    policy: Tuple[str, PolicyLanguage]
    ct: Ac17Ciphertext


@dataclass
class Ac17SecretKey:            # ac17/mod.rs:113
    k_0: List[bytes]
    k: List[Tuple[str, List[bytes]]]
    k_p: List[bytes]


@dataclass
class Ac17CpSecretKey:          # ac17/mod.rs:132
    attr: List[str]
    sk: Ac17SecretKey


def _pk_handle(pk: Ac17PublicKey):
    """Device-resident key material: fixed-base tables are built once per key and engine."""
    def build(e):
        h = e.ac17_pk_load(np.frombuffer(pk.to_bytes(), dtype=np.uint8))
        h.gt0 = e.gt_table(np.frombuffer(pk.e_gh_ka[0], dtype=np.uint8), 8)
        h.extra_handles = (h.gt0,)
        return h
    return HANDLES.get("ac17_pk", pk.to_bytes(), build)


def _msk_handle(msk: Ac17MasterKey):
    return HANDLES.get("ac17_msk", msk.to_bytes(), lambda e: e.ac17_msk_load(np.frombuffer(msk.to_bytes(), dtype=np.uint8)))


# ---------------------------------------------------------------------------------------------
def setup(rng: Rng = None) -> Tuple[Ac17PublicKey, Ac17MasterKey]:
    """ac17/mod.rs:141-188."""
    rng = rng or Rng()
    pk, msk = engine().ac17_setup(np.frombuffer(rng.frs(9), dtype=np.uint8))
    return Ac17PublicKey.from_bytes(pk), Ac17MasterKey.from_bytes(msk)


def cp_keygen(msk: Ac17MasterKey, attributes: List[str], rng: Rng = None) -> Ac17CpSecretKey:
    """ac17/mod.rs:191-264."""
    if len(attributes) == 0:
        raise RabeError("empty attributes!")
    return cp_keygen_batch(msk, attributes, 1, rng)[0]


def cp_keygen_batch(msk: Ac17MasterKey, attributes: List[str], count: int, rng: Rng = None) -> List[Ac17CpSecretKey]:
    if len(attributes) == 0:
        raise RabeError("empty attributes!")
    rng = rng or Rng()
    n = len(attributes)
    L = _lib.lib()
    h_attr, h_01 = (ctypes.c_uint8 * (192 * n))(), (ctypes.c_uint8 * 192)()
    _lib.check(L.rb_ac17_attr_hashes(_cstrs(attributes), n, h_attr, h_01), "rb_ac17_attr_hashes")
    rnd = np.frombuffer(rng.frs((n + 3) * count), dtype=np.uint8)
    k0, k, kp = engine().ac17_cp_keygen(_msk_handle(msk), np.frombuffer(bytes(h_attr), dtype=np.uint8),
                                        np.frombuffer(bytes(h_01), dtype=np.uint8), rnd, n)
    k0, k, kp = k0.tobytes(), k.tobytes(), kp.tobytes()
    keys = []
    for b in range(count):
        kb = k[192 * n * b:192 * n * (b + 1)]
        keys.append(Ac17CpSecretKey(
            attr=list(attributes),
            sk=Ac17SecretKey(k_0=[k0[384 * b + 128 * i:384 * b + 128 * (i + 1)] for i in range(3)],
                             k=[(a, [kb[192 * x + 64 * i:192 * x + 64 * (i + 1)] for i in range(3)]) for x, a in enumerate(attributes)],
                             k_p=[kp[192 * b + 64 * i:192 * b + 64 * (i + 1)] for i in range(3)])))
    return keys


def cp_encrypt(pk: Ac17PublicKey, policy: str, plaintext: bytes, language: PolicyLanguage, rng: Rng = None) -> Ac17CpCiphertext:
    """ac17/mod.rs:274-376."""
    return cp_encrypt_batch(pk, policy, [plaintext], language, rng)[0]


def cp_encrypt_batch(pk: Ac17PublicKey, policy: str, plaintexts: List[bytes], language: PolicyLanguage, rng: Rng = None):
    rng = rng or Rng()
    eng = engine()
    pol = Policy(policy, language)                  # Err(e) of parse() -> RabeError
    _, pi, _ = pol.msp()                            # AbePolicy::from_policy(...).unwrap()
    p = ctypes.c_void_p()
    _lib.check(eng.L.rb_ac17_msp_from_policy(eng.ctx, pol.ptr, ctypes.byref(p)), "rb_ac17_msp_from_policy")
    from ..engine import _Handle
    msp = _Handle(p, eng.L.rb_msp_free, eng)
    msp.n1 = len(pi)
    B = len(plaintexts)
    pkh = _pk_handle(pk)
    s = np.frombuffer(rng.frs(2 * B), dtype=np.uint8)            # `s` vector, :289-295
    msgs = eng.gt_pow_fixed(pkh.gt0, np.frombuffer(rng.frs(B), dtype=np.uint8))   # random Gt `msg`, :362
    c0, c, cp = [x.tobytes() for x in eng.ac17_cp_encrypt(pkh, msp, s, msgs)]
    msgs = msgs.tobytes()
    n1 = len(pi)
    out = []
    for b in range(B):
        cb = c[192 * n1 * b:192 * n1 * (b + 1)]
        out.append(Ac17CpCiphertext(
            policy=(policy, PolicyLanguage(language)),
            ct=Ac17Ciphertext(c_0=[c0[384 * b + 128 * i:384 * b + 128 * (i + 1)] for i in range(3)],
                              c=[(name, [cb[192 * x + 64 * i:192 * x + 64 * (i + 1)] for i in range(3)]) for x, name in enumerate(pi)],
                              c_p=cp[384 * b:384 * (b + 1)],
                              ct=encrypt_symmetric(msgs[384 * b:384 * (b + 1)], plaintexts[b], rng))))
    return out


def cp_decrypt(sk: Ac17CpSecretKey, ct: Ac17CpCiphertext) -> bytes:
    """ac17/mod.rs:385-430."""
    return cp_decrypt_batch(sk, [ct])[0]


def cp_decrypt_batch(sk: Ac17CpSecretKey, cts: List[Ac17CpCiphertext]) -> List[bytes]:
    """All ciphertexts must carry the same policy text (one pruned set per call)."""
    eng = engine()
    first = cts[0]
    if any(c.policy != first.policy for c in cts):
        raise RabeError("cp_decrypt_batch: ciphertexts with different policies must go in separate batches")
    pol = Policy(first.policy[0], first.policy[1])
    ct_names = [name for name, _ in first.ct.c]
    sk_names = [name for name, _ in sk.sk.k]
    n_ct, n_sk = len(ct_names), len(sk_names)
    matched = ctypes.c_int()
    cap = max(n_ct * max(n_sk, 1), 1) + n_ct + n_sk
    ct_idx, sk_idx = (ctypes.c_uint32 * cap)(), (ctypes.c_uint32 * cap)()
    nci, nsi = ctypes.c_uint32(), ctypes.c_uint32()
    st = eng.L.rb_ac17_decrypt_lists(pol.ptr, _cstrs(sk.attr), len(sk.attr), _cstrs(ct_names), n_ct, ctypes.byref(matched),
                                     ct_idx, cap, ctypes.byref(nci), sk_idx, cap, ctypes.byref(nsi))
    _lib.check(st, "rb_ac17_decrypt_lists")
    if not matched.value:
        raise RabeError("Error in cp_decrypt: attributes in SK do not match policy in CT.")
    # the key rows are matched by the names stored in sk.k (ac17/mod.rs:409); sk.attr drives pruning
    if sk_names != list(sk.attr):
        sk_rows = [i for cur, _ in pol.prune(sk.attr)[1] for i, n in enumerate(sk_names) if n == cur]
    else:
        sk_rows = list(sk_idx[:nsi.value])
    u8 = lambda b: np.frombuffer(b, dtype=np.uint8)
    k0 = u8(b"".join(sk.sk.k_0)); k = u8(b"".join(b"".join(v) for _, v in sk.sk.k)); kp = u8(b"".join(sk.sk.k_p))
    c0 = u8(b"".join(b"".join(c.ct.c_0) for c in cts))
    cc = u8(b"".join(b"".join(b"".join(v) for _, v in c.ct.c) for c in cts))
    cp = u8(b"".join(c.ct.c_p for c in cts))
    msgs = eng.ac17_cp_decrypt(k0, k, kp, c0, cc, cp, n_ct, list(ct_idx[:nci.value]), sk_rows).tobytes()
    return [decrypt_symmetric(msgs[384 * b:384 * (b + 1)], c.ct.ct) for b, c in enumerate(cts)]


# ============================================================================================ KP variant
@dataclass
class Ac17KpCiphertext:         # ac17/mod.rs:104
    attr: List[str]
    ct: Ac17Ciphertext


@dataclass
class Ac17KpSecretKey:          # ac17/mod.rs:122
    policy: Tuple[str, PolicyLanguage]
    sk: Ac17SecretKey


def _policy_hashes(pi, n2):
    from ..policy import sha3_hash_fr
    h_row = b"".join(sha3_hash_fr("%s%d%d" % (name, l, t)) for name in pi for l in range(3) for t in range(2))
    h_col = b"".join(sha3_hash_fr("0%d%d%d" % (j + 1, l, t)) for j in range(n2) for l in range(3) for t in range(2))
    return h_row, h_col


def kp_keygen(msk: Ac17MasterKey, policy: str, lang: PolicyLanguage, rng: Rng = None) -> Ac17KpSecretKey:
    """ac17/mod.rs:439-546."""
    rng = rng or Rng()
    pol = Policy(policy, lang)
    m, pi, c = pol.msp()
    n1 = len(pi)
    h_row, h_col = _policy_hashes(pi, c)
    rnd = np.frombuffer(rng.frs(2 + (c - 1) + n1), dtype=np.uint8)     # r0, r1, sigma'[..], sigma_attr[..]
    k0, k = engine().ac17_kp_keygen(_msk_handle(msk), np.array(m, dtype=np.int8).reshape(n1, c),
                                    np.frombuffer(h_row, dtype=np.uint8), np.frombuffer(h_col, dtype=np.uint8), rnd)
    k0, k = k0.tobytes(), k.tobytes()
    return Ac17KpSecretKey((policy, PolicyLanguage(lang)),
                           Ac17SecretKey(k_0=[k0[128 * i:128 * (i + 1)] for i in range(3)],
                                         k=[(name, [k[192 * x + 64 * i:192 * x + 64 * (i + 1)] for i in range(3)]) for x, name in enumerate(pi)],
                                         k_p=[]))


def kp_encrypt(pk: Ac17PublicKey, attributes: List[str], data: bytes, rng: Rng = None, _msg=None) -> Ac17KpCiphertext:
    """ac17/mod.rs:556-617 -- the CP row kernel with the degenerate policy (no columns to fold)."""
    rng = rng or Rng()
    eng = engine()
    n = len(attributes)
    from ..policy import sha3_hash_fr
    h_attr = b"".join(sha3_hash_fr("%s%d%d" % (a, l, t)) for a in attributes for l in range(3) for t in range(2))
    msp = eng.msp_load(np.zeros((n, 1), dtype=np.int8), np.frombuffer(h_attr, dtype=np.uint8), np.zeros(192, dtype=np.uint8))
    pkh = _pk_handle(pk)
    s = np.frombuffer(rng.frs(2), dtype=np.uint8)
    msg = _msg if _msg is not None else eng.gt_pow_fixed(pkh.gt0, np.frombuffer(rng.fr(), dtype=np.uint8)).tobytes()
    c0, c, cp = [x.tobytes() for x in eng.ac17_cp_encrypt(pkh, msp, s, np.frombuffer(msg, dtype=np.uint8))]
    return Ac17KpCiphertext(list(attributes),
                            Ac17Ciphertext(c_0=[c0[128 * i:128 * (i + 1)] for i in range(3)],
                                           c=[(a, [c[192 * x + 64 * i:192 * x + 64 * (i + 1)] for i in range(3)]) for x, a in enumerate(attributes)],
                                           c_p=cp, ct=encrypt_symmetric(msg, data, rng)))


def kp_decrypt_gt(sk: Ac17KpSecretKey, ct: Ac17KpCiphertext) -> bytes:
    eng = engine()
    pol = Policy(sk.policy[0], sk.policy[1])
    if not pol.satisfied(ct.attr):
        raise RabeError("Error in kp_decrypt: attributes in ct do not match policy in sk.")
    ok, lst = pol.prune(ct.attr)
    if not ok:
        raise RabeError("Error in kp_decrypt: pruned attributes in sk do not match policy in ct.")
    ct_names = [n for n, _ in ct.ct.c]
    sk_names = [n for n, _ in sk.sk.k]
    ct_idx = [i for cur, _ in lst for i, n in enumerate(ct_names) if n == cur]      # ac17/mod.rs:642-653
    sk_idx = [i for cur, _ in lst for i, n in enumerate(sk_names) if n == cur]
    u8 = lambda b: np.frombuffer(b, dtype=np.uint8)
    return eng.ac17_cp_decrypt(u8(b"".join(sk.sk.k_0)), u8(b"".join(b"".join(v) for _, v in sk.sk.k)), np.zeros(192, dtype=np.uint8),
                               u8(b"".join(ct.ct.c_0)), u8(b"".join(b"".join(v) for _, v in ct.ct.c)), u8(ct.ct.c_p), len(ct_names),
                               ct_idx, sk_idx).tobytes()


def kp_decrypt(sk: Ac17KpSecretKey, ct: Ac17KpCiphertext) -> bytes:
    """ac17/mod.rs:625-680."""
    return decrypt_symmetric(kp_decrypt_gt(sk, ct), ct.ct.ct)
