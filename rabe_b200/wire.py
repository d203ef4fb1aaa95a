"""Wire / on-disk formats of rabe's key and ciphertext structs.

Two layers, as in the reference:

* **struct framing** -- the `borsh` feature that `rabe-console` builds with (rabe-console/Cargo.toml:18-21, the
  `#[cfg_attr(feature = "borsh", derive(BorshSerialize, BorshDeserialize))]` on every struct, e.g.
  /root/reference/src/schemes/ac17/mod.rs:59-135): fields in declaration order; `Vec<T>` = u32 little-endian
  length + elements; `String` = u32 length + UTF-8; tuples = members in order; `PolicyLanguage` = one byte
  (JsonPolicy 0, HumanPolicy 1, pest/mod.rs:18-23); `Vec<u8>` = u32 length + bytes.  This part is fixed by the
  borsh specification and the struct definitions alone.
* **element encoding** -- how `rabe_bn` serialises Fr / G1 / G2 / Gt.  The crate is not in the reference tree and no
  reference test pins a serialised byte, so the codec is a PARAMETER here: `CanonicalCodec` writes
  this repository's canonical encodings (32-byte big-endian, affine, `include/rabe_b200.h`); a maintainer with a
  real rabe build supplies rabe_bn's byte layout as another `ElementCodec` and every struct below follows.
* **CLI envelope** -- `ser_enc` / `ser_dec` of /root/reference/rabe-console/src/mod.rs:1624-1683:
  `-----BEGIN X-----\\n` + hex(raw-deflate(borsh(struct))) + `\\n-----END X-----`, read back by taking the second line
  (`read_raw`, src/utils/file/mod.rs:70-76).

Host-side plumbing only (no group arithmetic): `rabe_b200.schemes.*` dataclasses in, bytes out, and back.
"""
import struct
import zlib
from typing import Any

from .policy import PolicyLanguage
from .schemes import ac17, aw11, bsw, lsw


class ElementCodec:
    """Byte layout of the four rabe_bn element types.  Sizes are fixed per codec; values are this repo's canonical bytes."""
    sizes = {"Fr": 32, "G1": 64, "G2": 128, "Gt": 384}

    def enc(self, kind: str, canonical: bytes) -> bytes:
        raise NotImplementedError

    def dec(self, kind: str, wire: bytes) -> bytes:
        raise NotImplementedError


class CanonicalCodec(ElementCodec):
    """The encodings of include/rabe_b200.h, unchanged."""

    def enc(self, kind, canonical):
        assert len(canonical) == self.sizes[kind], (kind, len(canonical))
        return bytes(canonical)

    def dec(self, kind, wire):
        return bytes(wire)


# ---- schema language: "Fr" | "G1" | "G2" | "Gt" | "String" | "Bytes" | "Lang" | ("vec", T) | ("tuple", T...) | ("struct", cls, [(field, T)...])
def _struct(cls, *fields):
    return ("struct", cls, list(fields))


POLICY = ("tuple", "String", "Lang")
AC17_CT = _struct(ac17.Ac17Ciphertext, ("c_0", ("vec", "G2")), ("c", ("vec", ("tuple", "String", ("vec", "G1")))), ("c_p", "Gt"), ("ct", "Bytes"))
AC17_SK = _struct(ac17.Ac17SecretKey, ("k_0", ("vec", "G2")), ("k", ("vec", ("tuple", "String", ("vec", "G1")))), ("k_p", ("vec", "G1")))
BSW_ATTR = _struct(bsw.CpAbeAttribute, ("string", "String"), ("g1", "G1"), ("g2", "G2"))
SCHEMAS = {
    # ac17/mod.rs:59-135
    ac17.Ac17PublicKey: _struct(ac17.Ac17PublicKey, ("g", "G1"), ("h_a", ("vec", "G2")), ("e_gh_ka", ("vec", "Gt"))),
    ac17.Ac17MasterKey: _struct(ac17.Ac17MasterKey, ("g", "G1"), ("h", "G2"), ("g_k", ("vec", "G1")), ("a", ("vec", "Fr")), ("b", ("vec", "Fr"))),
    ac17.Ac17Ciphertext: AC17_CT,
    ac17.Ac17CpCiphertext: _struct(ac17.Ac17CpCiphertext, ("policy", POLICY), ("ct", AC17_CT)),
    ac17.Ac17SecretKey: AC17_SK,
    ac17.Ac17CpSecretKey: _struct(ac17.Ac17CpSecretKey, ("attr", ("vec", "String")), ("sk", AC17_SK)),
    # bsw/mod.rs:40-90
    bsw.CpAbePublicKey: _struct(bsw.CpAbePublicKey, ("g1", "G1"), ("g2", "G2"), ("h", "G1"), ("f", "G2"), ("e_gg_alpha", "Gt")),
    bsw.CpAbeMasterKey: _struct(bsw.CpAbeMasterKey, ("beta", "Fr"), ("g2_alpha", "G2")),
    bsw.CpAbeAttribute: BSW_ATTR,
    bsw.CpAbeCiphertext: _struct(bsw.CpAbeCiphertext, ("policy", POLICY), ("c", "G1"), ("c_p", "Gt"), ("c_y", ("vec", BSW_ATTR)), ("data", "Bytes")),
    bsw.CpAbeSecretKey: _struct(bsw.CpAbeSecretKey, ("d", "G2"), ("d_j", ("vec", BSW_ATTR))),
    # lsw/mod.rs:41-83
    lsw.KpAbePublicKey: _struct(lsw.KpAbePublicKey, ("g1", "G1"), ("g2", "G2"), ("g1_b", "G1"), ("g1_b2", "G1"), ("h_b", "G1"), ("e_gg_alpha", "Gt")),
    lsw.KpAbeMasterKey: _struct(lsw.KpAbeMasterKey, ("alpha1", "Fr"), ("alpha2", "Fr"), ("b", "Fr"), ("h_g1", "G1"), ("h_g2", "G2")),
    lsw.KpAbeSecretKey: _struct(lsw.KpAbeSecretKey, ("policy", POLICY), ("dj", ("vec", ("tuple", "String", "G1", "G2", "G1", "G1", "G1")))),
    lsw.KpAbeCiphertext: _struct(lsw.KpAbeCiphertext, ("e1", "Gt"), ("e2", "G2"), ("ej", ("vec", ("tuple", "String", "G1", "G1", "G1"))), ("ct", "Bytes")),
    # aw11/mod.rs:46-92
    aw11.Aw11GlobalKey: _struct(aw11.Aw11GlobalKey, ("g1", "G1"), ("g2", "G2")),
    aw11.Aw11PublicKey: _struct(aw11.Aw11PublicKey, ("attr", ("vec", ("tuple", "String", "Gt", "G2")))),
    aw11.Aw11MasterKey: _struct(aw11.Aw11MasterKey, ("attr", ("vec", ("tuple", "String", "Fr", "Fr")))),
    aw11.Aw11Ciphertext: _struct(aw11.Aw11Ciphertext, ("policy", POLICY), ("c_0", "Gt"), ("c", ("vec", ("tuple", "String", "Gt", "G2", "G2"))), ("ct", "Bytes")),
    aw11.Aw11SecretKey: _struct(aw11.Aw11SecretKey, ("gid", "String"), ("attr", ("vec", ("tuple", "String", "G1")))),
}
for _name in ("Ac17KpCiphertext", "Ac17KpSecretKey"):
    _cls = getattr(ac17, _name, None)
    if _cls is not None:
        SCHEMAS[_cls] = (_struct(_cls, ("attr", ("vec", "String")), ("ct", AC17_CT)) if _name == "Ac17KpCiphertext"
                         else _struct(_cls, ("policy", POLICY), ("sk", AC17_SK)))

# header / footer tags of the CLI envelope (rabe-console/src/mod.rs:82-99)
TAGS = {"GP": "GP", "SK": "SK", "MSK": "MSK", "PK": "PK", "CT": "CT", "SKA": "SAK", "PKA": "PAK", "AU_PK": "PAUK", "AU_SK": "SAUK"}


class WireError(ValueError):
    pass


def _put(t: Any, v: Any, codec: ElementCodec, out: bytearray):
    if isinstance(t, str):
        if t in codec.sizes:
            out += codec.enc(t, v)
        elif t == "String":
            b = v.encode("utf-8"); out += struct.pack("<I", len(b)) + b
        elif t == "Bytes":
            out += struct.pack("<I", len(v)) + bytes(v)
        elif t == "Lang":
            out.append(int(PolicyLanguage(v)))
        else:
            raise WireError("unknown type " + t)
    elif t[0] == "vec":
        out += struct.pack("<I", len(v))
        for x in v:
            _put(t[1], x, codec, out)
    elif t[0] == "tuple":
        if len(v) != len(t) - 1:
            raise WireError("tuple arity")
        for tt, x in zip(t[1:], v):
            _put(tt, x, codec, out)
    else:                                        # struct
        for name, tt in t[2]:
            _put(tt, getattr(v, name), codec, out)


def _get(t: Any, buf: memoryview, pos: int, codec: ElementCodec):
    def need(n):
        if pos + n > len(buf):
            raise WireError("unexpected end of input")       # borsh: "Unexpected length of input"
    if isinstance(t, str):
        if t in codec.sizes:
            n = len(codec.enc(t, b"\0" * codec.sizes[t])); need(n)
            return codec.dec(t, bytes(buf[pos:pos + n])), pos + n
        if t in ("String", "Bytes"):
            need(4); (n,) = struct.unpack_from("<I", buf, pos); pos += 4
            if pos + n > len(buf):
                raise WireError("unexpected end of input")
            raw = bytes(buf[pos:pos + n])
            return (raw.decode("utf-8") if t == "String" else raw), pos + n
        if t == "Lang":
            need(1)
            if buf[pos] > 1:
                raise WireError("invalid PolicyLanguage variant")
            return PolicyLanguage(buf[pos]), pos + 1
        raise WireError("unknown type " + t)
    if t[0] == "vec":
        need(4); (n,) = struct.unpack_from("<I", buf, pos); pos += 4
        items = []
        for _ in range(n):
            x, pos = _get(t[1], buf, pos, codec); items.append(x)
        return items, pos
    if t[0] == "tuple":
        items = []
        for tt in t[1:]:
            x, pos = _get(tt, buf, pos, codec); items.append(x)
        return tuple(items), pos
    vals = []
    for _, tt in t[2]:
        x, pos = _get(tt, buf, pos, codec); vals.append(x)
    return t[1](*vals), pos


def to_borsh(obj, codec: ElementCodec = None) -> bytes:
    """borsh bytes of a rabe_b200.schemes dataclass (struct framing of the reference, element bytes of `codec`)."""
    out = bytearray()
    _put(SCHEMAS[type(obj)], obj, codec or CanonicalCodec(), out)
    return bytes(out)


def from_borsh(cls, data: bytes, codec: ElementCodec = None):
    obj, pos = _get(SCHEMAS[cls], memoryview(bytes(data)), 0, codec or CanonicalCodec())
    if pos != len(data):
        raise WireError("trailing bytes")                     # borsh try_from_slice: "Not all bytes read"
    return obj


def ser_enc(obj, tag: str, codec: ElementCodec = None) -> str:
    """rabe-console `ser_enc` (mod.rs:1624-1636): header + hex(deflate(borsh)) + footer.  tag: "PK", "MSK", "SK", "CT", ..."""
    raw = zlib.compressobj(9, zlib.DEFLATED, -15)
    body = raw.compress(to_borsh(obj, codec)) + raw.flush()
    return "-----BEGIN %s-----\n%s\n-----END %s-----" % (tag, body.hex(), tag)


def ser_dec(cls, text: str, codec: ElementCodec = None):
    """rabe-console `ser_dec` (mod.rs:1651-1683): the second line (`read_raw`), hex -> inflate -> borsh."""
    lines = text.splitlines()
    if len(lines) < 2:
        raise WireError("read_raw: no second line")
    try:
        data = zlib.decompress(bytes.fromhex(lines[1]), -15)
    except (ValueError, zlib.error) as e:
        raise WireError("inflate_bytes: %s" % e)
    return from_borsh(cls, data, codec)
