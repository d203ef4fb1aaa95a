"""KEM tail on the device (rb_kem_encrypt_batch / rb_kem_decrypt_batch) against rabe's encrypt_symmetric /
decrypt_symmetric as restated on the host with hashlib + `cryptography` (aes/mod.rs:10-55): SHA3-256 of the Gt
encoding as the key, AES-256-GCM, nonce | ciphertext | tag -- bit-exact, every payload length from 0 to 70 bytes
(partial blocks, block boundaries) plus a few KB, and forged items rejected."""
import hashlib
import random

import numpy as np
import pytest

from rb_testutil import u8

pytestmark = pytest.mark.gpu


def test_kem_matches_aesgcm_and_rejects_forgeries(engine):
    from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    rng = random.Random(71)
    rb = lambda n: bytes(rng.randrange(256) for _ in range(n))
    lengths = list(range(0, 71)) + [127, 128, 129, 1000, 4096, 5001]
    B = len(lengths)
    gts = [rb(384) for _ in range(B)]                      # the KDF hashes the bytes; they need not be group elements
    nonces = [rb(12) for _ in range(B)]
    payloads = [rb(n) for n in lengths]
    blobs = engine.kem_encrypt(u8(b"".join(gts)), u8(b"".join(nonces)), payloads)
    for gt, nonce, data, blob in zip(gts, nonces, payloads, blobs):
        want = nonce + AESGCM(hashlib.sha3_256(gt).digest()).encrypt(nonce, data, None)
        assert blob == want, len(data)
    assert engine.kem_decrypt(u8(b"".join(gts)), blobs) == payloads
    # forged: a flipped ciphertext bit, a flipped tag bit, the wrong Gt (wrong key), a truncated blob
    bad = list(blobs)
    bad[5] = bad[5][:13] + bytes([bad[5][13] ^ 1]) + bad[5][14:]
    bad[9] = bad[9][:-1] + bytes([bad[9][-1] ^ 0x80])
    bad[20] = bad[20][:20]
    gts2 = list(gts); gts2[30] = rb(384)
    got = engine.kem_decrypt(u8(b"".join(gts2)), bad)
    for b in range(B):
        assert got[b] == (None if b in (5, 9, 20, 30) else payloads[b]), b
