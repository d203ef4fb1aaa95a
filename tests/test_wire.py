"""Wire formats (rabe_b200/wire.py): the borsh struct framing of rabe's key / ciphertext structs (field order, u32
length prefixes, String / tuple / enum layout -- fixed by the reference's struct definitions and the borsh
specification), the element codec as a parameter, and the rabe-console file envelope.  CPU only."""
import os
import random
import struct
import zlib

import pytest

from rabe_b200 import wire
from rabe_b200.policy import PolicyLanguage
from rabe_b200.schemes import ac17, aw11, bsw, lsw

rng = random.Random(61)
rb = lambda n: bytes(rng.randrange(256) for _ in range(n))
FR, G1, G2, GT = (lambda: rb(32)), (lambda: rb(64)), (lambda: rb(128)), (lambda: rb(384))


def samples():
    pol = ('"A" and "B"', PolicyLanguage.HumanPolicy)
    ct = ac17.Ac17Ciphertext([G2(), G2(), G2()], [("A", [G1(), G1(), G1()]), ("B", [G1(), G1(), G1()])], GT(), rb(71))
    sk = ac17.Ac17SecretKey([G2(), G2(), G2()], [("attr1", [G1(), G1(), G1()])], [G1(), G1(), G1()])
    attr = lambda s: bsw.CpAbeAttribute(s, G1(), G2())
    return [
        ac17.Ac17PublicKey(G1(), [G2(), G2(), G2()], [GT(), GT()]),
        ac17.Ac17MasterKey(G1(), G2(), [G1(), G1(), G1()], [FR(), FR()], [FR(), FR()]),
        ct, ac17.Ac17CpCiphertext(pol, ct), sk, ac17.Ac17CpSecretKey(["attr1", "x"], sk),
        bsw.CpAbePublicKey(G1(), G2(), G1(), G2(), GT()), bsw.CpAbeMasterKey(FR(), G2()),
        bsw.CpAbeCiphertext(('{"name": "A"}', PolicyLanguage.JsonPolicy), G1(), GT(), [attr("A_10"), attr("B_20")], rb(40)),
        bsw.CpAbeSecretKey(G2(), [attr("A"), attr("B"), attr("C")]),
        lsw.KpAbePublicKey(G1(), G2(), G1(), G1(), G1(), GT()), lsw.KpAbeMasterKey(FR(), FR(), FR(), G1(), G2()),
        lsw.KpAbeSecretKey(pol, [("A_2", G1(), G2(), G1(), G1(), G1())]), lsw.KpAbeCiphertext(GT(), G2(), [("A", G1(), G1(), G1()), ("B", G1(), G1(), G1())], rb(33)),
        aw11.Aw11GlobalKey(G1(), G2()), aw11.Aw11PublicKey([("A", GT(), G2())]), aw11.Aw11MasterKey([("A", FR(), FR()), ("B", FR(), FR())]),
        aw11.Aw11Ciphertext(pol, GT(), [("A_2", GT(), G2(), G2())], rb(5)), aw11.Aw11SecretKey("bob", [("A", G1()), ("B", G1())]),
    ]


def test_every_struct_round_trips_through_borsh_and_the_cli_envelope():
    for obj in samples():
        data = wire.to_borsh(obj)
        assert wire.from_borsh(type(obj), data) == obj
        text = wire.ser_enc(obj, "CT")
        assert text.startswith("-----BEGIN CT-----\n") and text.endswith("\n-----END CT-----") and len(text.splitlines()) == 3
        assert wire.ser_dec(type(obj), text) == obj
        with pytest.raises(wire.WireError):
            wire.from_borsh(type(obj), data[:-1])
        with pytest.raises(wire.WireError):
            wire.from_borsh(type(obj), data + b"\0")


def test_borsh_layout_by_hand():
    """field order, little-endian u32 prefixes, tuple / String / enum layout (borsh spec + aw11/mod.rs:86-92, ac17/mod.rs:95-99)"""
    g = G1()
    sk = aw11.Aw11SecretKey("bob", [("A", g)])
    assert wire.to_borsh(sk) == struct.pack("<I", 3) + b"bob" + struct.pack("<I", 1) + struct.pack("<I", 1) + b"A" + g
    ct = ac17.Ac17Ciphertext([], [], b"\x11" * 384, b"xyz")
    cp = ac17.Ac17CpCiphertext(("A", PolicyLanguage.HumanPolicy), ct)
    want = struct.pack("<I", 1) + b"A" + b"\x01" + struct.pack("<I", 0) + struct.pack("<I", 0) + b"\x11" * 384 + struct.pack("<I", 3) + b"xyz"
    assert wire.to_borsh(cp) == want
    with pytest.raises(wire.WireError):
        wire.from_borsh(ac17.Ac17CpCiphertext, want[:5] + b"\x02" + want[6:])          # PolicyLanguage has two variants


def test_envelope_is_hex_of_raw_deflate_and_reads_the_second_line():
    obj = samples()[0]
    text = wire.ser_enc(obj, "PK")
    assert zlib.decompress(bytes.fromhex(text.splitlines()[1]), -15) == wire.to_borsh(obj)
    assert wire.ser_dec(type(obj), "garbage first line\n" + text.splitlines()[1] + "\nanything") == obj     # read_raw: lines().nth(1)
    with pytest.raises(wire.WireError):
        wire.ser_dec(type(obj), "one line only")


def test_element_codec_is_a_parameter():
    """another element encoding (here: little-endian limbs, a stand-in for whatever rabe_bn writes) changes the element
    bytes and nothing else of the framing"""
    class LittleEndian(wire.ElementCodec):
        def enc(self, kind, canonical):
            return b"".join(canonical[i:i + 32][::-1] for i in range(0, len(canonical), 32))
        dec = enc
    obj = samples()[18]                                  # Aw11SecretKey
    a, b = wire.to_borsh(obj), wire.to_borsh(obj, LittleEndian())
    assert len(a) == len(b) and a != b and a[:7] == b[:7]
    assert wire.from_borsh(type(obj), b, LittleEndian()) == obj
