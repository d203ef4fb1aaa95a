"""Shared host-side plumbing of the scheme mirrors: randomness source, the KEM tail
(/root/reference/src/utils/aes/mod.rs), the process-wide engine and small byte helpers."""
import hashlib
import os

import numpy as np

from ..engine import Engine
from ..error import RabeError

R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G1_GEN = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
G2_GEN = b"".join(x.to_bytes(32, "big") for x in (
    10857046999023057135944570762232829481370756359578518086990519993285655852781,
    11559732032986387107991004021392285783925812861821192530917403151452391805634,
    8495653923123431417604973247489272438418190587263600148770280649306958101930,
    4082367875863433681332203403145435568316851327593401208105741076214120093531))
FR_ONE = (1).to_bytes(32, "big")
FR_MINUS_ONE = (R_ORDER - 1).to_bytes(32, "big")


class Rng:
    """Source of the scalars rabe draws with `rng.gen()`.  seed=None -> os.urandom; a seed gives a
    reproducible SHA3-512 counter stream; `values` (iterable of ints) replays an explicit sequence."""

    def __init__(self, seed=None, values=None):
        self._ctr, self._seed = 0, (None if seed is None else str(seed).encode())
        self._values = iter(values) if values is not None else None

    def _bytes(self, n):
        if self._seed is None:
            return os.urandom(n)
        out = b""
        while len(out) < n:
            out += hashlib.sha3_512(self._seed + self._ctr.to_bytes(8, "big")).digest()
            self._ctr += 1
        return out[:n]

    def fr(self) -> bytes:
        if self._values is not None:
            return (int(next(self._values)) % R_ORDER).to_bytes(32, "big")
        return (int.from_bytes(self._bytes(64), "big") % R_ORDER).to_bytes(32, "big")

    def frs(self, n) -> bytes:
        return b"".join(self.fr() for _ in range(n))

    def nonce(self) -> bytes:
        # a seed gives a reproducible nonce (tests); a replayed `values` stream carries no nonce
        # material, so the nonce is fresh -- never a constant (AES-GCM key/nonce reuse otherwise)
        return self._bytes(12) if self._seed is not None else os.urandom(12)


_ENGINE = None


def engine(device=None) -> Engine:
    """Process-wide engine (one rb_ctx); created on first use on cuda:`device` (default 0)."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(0 if device is None else device)
    return _ENGINE


def set_engine(e: Engine):
    """Switch the process-wide engine; device handles cached for the previous one are dropped first
    (they belong to its rb_ctx / device / stream)."""
    global _ENGINE
    if e is not _ENGINE:
        HANDLES.clear()
    _ENGINE = e


def u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8)


def chunks(b, size):
    b = bytes(b)
    return [b[i:i + size] for i in range(0, len(b), size)]


def _kdf(gt: bytes) -> bytes:                       # aes/mod.rs:47-55
    return hashlib.sha3_256(gt).digest()


def encrypt_symmetric(msg_gt: bytes, data: bytes, rng: Rng) -> bytes:      # aes/mod.rs:10-27
    from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    nonce = rng.nonce()
    return nonce + AESGCM(_kdf(msg_gt)).encrypt(nonce, data, None)


def decrypt_symmetric(msg_gt: bytes, nonce_ct: bytes) -> bytes:            # aes/mod.rs:29-45
    from cryptography.exceptions import InvalidTag
    from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    try:
        return AESGCM(_kdf(msg_gt)).decrypt(nonce_ct[:12], nonce_ct[12:], None)
    except InvalidTag:
        raise RabeError("decryption error: aead::Error")


class HandleCache:
    """Device-resident key material (fixed-base tables, loaded keys), built once per key and engine.
    Keyed by (engine, kind, SHA-256 of the key bytes) -- never by the raw (master) key bytes --, bounded
    (least recently used handles are freed) and emptied by set_engine() / clear()."""

    def __init__(self, capacity=32):
        from collections import OrderedDict
        self._d, self.capacity = OrderedDict(), capacity

    def get(self, kind, key_bytes, build):
        key = (id(engine()), kind, hashlib.sha256(bytes(key_bytes)).digest())
        h = self._d.get(key)
        if h is None:
            h = build(engine())
            self._d[key] = h
            while len(self._d) > self.capacity:
                _, old = self._d.popitem(last=False)
                _close(old)
        else:
            self._d.move_to_end(key)
        return h

    def clear(self):
        while self._d:
            _, old = self._d.popitem(last=False)
            _close(old)


def _close(h):
    for sub in getattr(h, "extra_handles", ()):
        sub.close()
    h.close()


HANDLES = HandleCache()


class TableCache:
    """Fixed-base tables keyed by the base's bytes (HandleCache entries of kind g1 / g2 / gt)."""

    def get(self, kind, base: bytes, w):
        build = lambda e: {"g1": e.g1_table, "g2": e.g2_table, "gt": e.gt_table}[kind](u8(base), w)
        return HANDLES.get((kind, w), base, build)


TABLES = TableCache()
