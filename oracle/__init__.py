"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/liboracle.so).

The oracle restates, on the CPU, the operation sequence of the reference for the hot path
(/root/reference/src/schemes/{ac17,bsw,lsw,aw11}/mod.rs over the external crate rabe-bn 0.4.23).
It is the checker for the CUDA path and the timed CPU baseline; it is never the product:
only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() import it.

PARITY UNPINNED against rabe itself (no Rust toolchain, rabe-bn source absent, no reference test
pins a group element); pinned instead against oracle/pyref.py and algebraic known answers.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

G1_BYTES, G2_BYTES, GT_BYTES, FR_BYTES = 64, 128, 384, 32


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            so = build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _buf(n):
    return (ctypes.c_uint8 * n)()


def _b(x):
    return (ctypes.c_uint8 * len(x)).from_buffer_copy(bytes(x))


def _strs(names):
    arr = (ctypes.c_char_p * len(names))()
    arr[:] = [n.encode() if isinstance(n, str) else n for n in names]
    return arr


def fr_bytes(x: int) -> bytes:
    return int(x).to_bytes(32, "big")


def _chk(rc, what):
    if rc != 0:
        raise ValueError(f"oracle {what} failed rc={rc}")


# ------------------------------------------------------------------ primitives
def ops_reset():
    lib().orc_ops_reset()


def ops_get():
    out = (ctypes.c_uint64 * 2)()
    lib().orc_ops_get(out)
    return {"fp_mul": out[0], "fr_mul": out[1]}


def constants():
    out = (ctypes.c_uint64 * 26)()
    lib().orc_constants(out)
    w = list(out)
    j = lambda ws: sum(x << (64 * i) for i, x in enumerate(ws))
    return {"p": j(w[0:4]), "R_p": j(w[4:8]), "R2_p": j(w[8:12]), "inv_p": w[12],
            "r": j(w[13:17]), "R_r": j(w[17:21]), "R2_r": j(w[21:25]), "inv_r": w[25]}


def g1_generator():
    o = _buf(64); lib().orc_g1_generator(o); return bytes(o)


def g2_generator():
    o = _buf(128); lib().orc_g2_generator(o); return bytes(o)


def g1_mul(base, k):
    o = _buf(64); _chk(lib().orc_g1_mul(_b(base), _b(k), o), "g1_mul"); return bytes(o)


def g2_mul(base, k):
    o = _buf(128); _chk(lib().orc_g2_mul(_b(base), _b(k), o), "g2_mul"); return bytes(o)


def g1_add(a, b):
    o = _buf(64); _chk(lib().orc_g1_add(_b(a), _b(b), o), "g1_add"); return bytes(o)


def g2_add(a, b):
    o = _buf(128); _chk(lib().orc_g2_add(_b(a), _b(b), o), "g2_add"); return bytes(o)


def g1_neg(a):
    o = _buf(64); _chk(lib().orc_g1_neg(_b(a), o), "g1_neg"); return bytes(o)


def g1_check(a):
    return lib().orc_g1_check(_b(a)) == 0


def g2_check(a):
    """decode-time check of the lineage's G2: on the twist AND [r]Q == O"""
    return lib().orc_g2_check(_b(a)) == 0


def g2_on_curve(a):
    return lib().orc_g2_on_curve(_b(a)) == 0


def pairing(p, q):
    o = _buf(384); _chk(lib().orc_pairing(_b(p), _b(q), o), "pairing"); return bytes(o)


def final_exp(f):
    o = _buf(384); _chk(lib().orc_final_exp(_b(f), o), "final_exp"); return bytes(o)


def gt_pow(a, k):
    o = _buf(384); _chk(lib().orc_gt_pow(_b(a), _b(k), o), "gt_pow"); return bytes(o)


def gt_mul(a, b):
    o = _buf(384); _chk(lib().orc_gt_mul(_b(a), _b(b), o), "gt_mul"); return bytes(o)


def gt_inverse(a):
    o = _buf(384); _chk(lib().orc_gt_inverse(_b(a), o), "gt_inverse"); return bytes(o)


def gt_cyclotomic_sqr(a):
    o = _buf(384); _chk(lib().orc_gt_cyclotomic_sqr(_b(a), o), "gt_cyclotomic_sqr"); return bytes(o)


def gt_frobenius(a, j):
    o = _buf(384); _chk(lib().orc_gt_frobenius(_b(a), j, o), "gt_frobenius"); return bytes(o)


GT_ONE = (b"\x00" * 31 + b"\x01") + b"\x00" * (384 - 32)

_FR_OPS = {"add": 0, "sub": 1, "mul": 2, "inverse": 3, "neg": 4, "pow": 5}


def fr_op(op, a, b=None):
    o = _buf(32)
    lib().orc_fr_op(_FR_OPS[op], _b(a), _b(b) if b is not None else None, o)
    return bytes(o)


def fq_op(op, a, b=None):
    o = _buf(32)
    lib().orc_fq_op(_FR_OPS[op], _b(a), _b(b) if b is not None else None, o)
    return bytes(o)


def sha3_256(data: bytes):
    o = _buf(32); lib().orc_sha3_256(_b(data) if data else None, ctypes.c_size_t(len(data)), o); return bytes(o)


def sha3_fr(s: str):
    o = _buf(32); e = s.encode(); lib().orc_sha3_fr(e, ctypes.c_size_t(len(e)), o); return bytes(o)


# ------------------------------------------------------------------ AC17 (reference sequence)
def ac17_setup(rnd: bytes):
    """rnd = 9 Fr (rho_g, rho_h, a0, b0, a1, b1, k0, k1, k2). Returns (pk[1216], msk[512])."""
    assert len(rnd) == 9 * 32
    pk, msk = _buf(1216), _buf(512)
    _chk(lib().orc_ac17_setup(_b(rnd), pk, msk), "ac17_setup")
    return bytes(pk), bytes(msk)


def ac17_cp_keygen(msk, attrs, rnd):
    """rnd = (n+3) Fr: r0, r1, sigma_attr[n], sigma. Returns (k_0[384], k[n*192], k_p[192])."""
    n = len(attrs)
    assert len(rnd) == (n + 3) * 32
    k0, k, kp = _buf(384), _buf(192 * max(n, 1)), _buf(192)
    _chk(lib().orc_ac17_cp_keygen(_b(msk), n, _strs(attrs), _b(rnd), k0, k, kp), "ac17_cp_keygen")
    return bytes(k0), bytes(k)[:192 * n], bytes(kp)


def ac17_cp_encrypt(pk, m_rows, pi, rnd, msg):
    """m_rows = list of n1 rows of n2 entries in {-1,0,1}; pi = row labels; rnd = s0|s1; msg Gt."""
    n1, n2 = len(m_rows), len(m_rows[0])
    flat = (ctypes.c_int8 * (n1 * n2))(*[v for row in m_rows for v in row])
    c0, c, cp = _buf(384), _buf(192 * n1), _buf(384)
    _chk(lib().orc_ac17_cp_encrypt(_b(pk), n1, n2, flat, _strs(pi), _b(rnd), _b(msg), c0, c, cp), "ac17_cp_encrypt")
    return bytes(c0), bytes(c), bytes(cp)


def ac17_cp_decrypt(pruned_names, ct_names, c0, c, cp, sk_names, k0, k, kp):
    out = _buf(384)
    _chk(lib().orc_ac17_cp_decrypt(len(pruned_names), _strs(pruned_names),
                                   len(ct_names), _strs(ct_names), _b(c0), _b(c), _b(cp),
                                   len(sk_names), _strs(sk_names), _b(k0), _b(k), _b(kp), out), "ac17_cp_decrypt")
    return bytes(out)


# ------------------------------------------------------------------ AC17, restructured algorithm (oracle/ac17_fast.cpp)
class Ac17Fast:
    """The algebra of the CUDA path on the CPU (fixed-base window tables, folded policy scalars, one final
    exponentiation per decryption) -- byte-identical outputs to the reference-sequence functions above; bench.py times
    both to split the GPU speed-up into its algorithmic and its hardware part."""

    def __init__(self, pk, m_rows, pi):
        L = lib()
        L.orc_ac17_fast_pk_new.restype = ctypes.c_void_p
        L.orc_ac17_fast_msp_new.restype = ctypes.c_void_p
        self.n1, n2 = len(m_rows), len(m_rows[0])
        flat = (ctypes.c_int8 * (self.n1 * n2))(*[v for row in m_rows for v in row])
        self.pk = ctypes.c_void_p(L.orc_ac17_fast_pk_new(_b(pk)))
        self.msp = ctypes.c_void_p(L.orc_ac17_fast_msp_new(self.n1, n2, flat, _strs(pi)))
        if not self.pk or not self.msp:
            raise ValueError("oracle ac17_fast: bad public key")

    def encrypt(self, rnd, msg):
        c0, c, cp = _buf(384), _buf(192 * self.n1), _buf(384)
        _chk(lib().orc_ac17_cp_encrypt_fast(self.pk, self.msp, _b(rnd), _b(msg), c0, c, cp), "ac17_cp_encrypt_fast")
        return bytes(c0), bytes(c), bytes(cp)

    @staticmethod
    def decrypt(ct_idx, sk_idx, c0, c, cp, k0, k, kp):
        out = _buf(384)
        ci = (ctypes.c_uint32 * max(1, len(ct_idx)))(*ct_idx); si = (ctypes.c_uint32 * max(1, len(sk_idx)))(*sk_idx)
        _chk(lib().orc_ac17_cp_decrypt_fast(len(ct_idx), ci, len(sk_idx), si, _b(c0), _b(c), _b(cp), _b(k0), _b(k), _b(kp), out), "ac17_cp_decrypt_fast")
        return bytes(out)

    def close(self):
        if self.pk:
            lib().orc_ac17_fast_pk_free(self.pk); lib().orc_ac17_fast_msp_free(self.msp)
            self.pk = self.msp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
