"""Thin Python front-end of the C ABI: one `Engine` = one rb_ctx (one GPU, one stream).

Buffers are numpy uint8 arrays / bytes (host; staged by the library) or torch CUDA uint8 tensors
(device resident; calls only enqueue work).  Outputs are created in the same residency as the
first data input.  All arithmetic happens in librabe_b200.so on the GPU.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check

FR, G1, G2, GT = 32, 64, 128, 384

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_cuda_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def _as_buf(x):
    """-> (pointer:int, keepalive, is_device)"""
    if x is None:
        return None, None, False
    if _is_cuda_tensor(x):
        assert x.dtype in (torch.uint8, torch.int8, torch.int32, torch.uint32) and x.is_contiguous()
        return x.data_ptr(), x, True
    if torch is not None and isinstance(x, torch.Tensor):
        x = x.numpy()
    if isinstance(x, (bytes, bytearray, memoryview)):
        x = np.frombuffer(bytes(x), dtype=np.uint8)
    x = np.ascontiguousarray(x)
    return x.ctypes.data, x, False


def _settle(*xs):
    """Handle-building calls are synchronous and run on the context's stream: device-resident inputs written on
    torch's current stream are made complete first."""
    if any(_is_cuda_tensor(x) for x in xs):
        torch.cuda.current_stream().synchronize()


class _Handle:
    def __init__(self, ptr, free, owner):
        self.ptr, self._free, self._owner = ptr, free, owner

    def close(self):
        if self.ptr:
            self._free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, device=0):
        self.L = _lib.lib()
        p = ctypes.c_void_p()
        check(self.L.rb_ctx_create(int(device), ctypes.byref(p)), "rb_ctx_create")
        self.ctx = p
        self.device = int(device)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.rb_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def use_torch_stream(self):
        """Enqueue on torch's current CUDA stream (so torch.cuda.Event timing brackets the kernels)."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        check(self.L.rb_ctx_set_stream(self.ctx, ctypes.c_void_p(s)), "rb_ctx_set_stream")
        self._ext = None

    def _ctx_stream(self):
        """torch view of the stream the context launches on (its own non-blocking stream unless use_torch_stream())."""
        ptr = self.L.rb_ctx_get_stream(self.ctx) or 0
        if getattr(self, "_ext", None) is None or self._ext.cuda_stream != ptr:
            self._ext = torch.cuda.ExternalStream(ptr, device=torch.device("cuda", self.device))
        return self._ext

    def sync(self):
        check(self.L.rb_ctx_sync(self.ctx), "rb_ctx_sync")

    def status(self):
        check(self.L.rb_ctx_status(self.ctx), "rb_ctx_status")

    def launch_count(self):
        return int(self.L.rb_ctx_launch_count(self.ctx))

    def set_g2_subgroup_check(self, enable=True):
        """Default on: caller-supplied G2 points are tested for subgroup membership (as rabe_bn does when it
        decodes them).  Off = the caller vouches for them (already validated / produced by this library)."""
        check(self.L.rb_ctx_set_g2_subgroup_check(self.ctx, 1 if enable else 0), "rb_ctx_set_g2_subgroup_check")

    PAIRING_AUTO, PAIRING_THROUGHPUT, PAIRING_LATENCY = 0, 1, 2

    def set_pairing_layout(self, mode):
        """AUTO (default): six-lane latency kernels unless another context of this GPU has a pairing batch in flight;
        THROUGHPUT: two-lane kernels; LATENCY: six-lane kernels.  Results are identical."""
        check(self.L.rb_ctx_set_pairing_layout(self.ctx, int(mode)), "rb_ctx_set_pairing_layout")

    def set_async(self, enable=True):
        """Host-buffer calls enqueue and return; outputs / status arrive with sync() / status()."""
        check(self.L.rb_ctx_set_async(self.ctx, 1 if enable else 0), "rb_ctx_set_async")

    def g2_check(self, q) -> bool:
        """True iff every G2 point of `q` is canonical, on the twist and in the order-r subgroup."""
        _settle(q)
        ptr, keep, _ = _as_buf(q)
        st = self.L.rb_g2_check_batch(self.ctx, ctypes.c_void_p(ptr), _nbytes(q) // G2)
        if st == _lib.RB_ENOTMEMBER:
            return False
        check(st, "rb_g2_check_batch")
        return True

    def profile(self, enable=True):
        check(self.L.rb_ctx_profile(self.ctx, 1 if enable else 0), "rb_ctx_profile")

    def profile_report(self):
        import json
        need = ctypes.c_size_t()
        check(self.L.rb_ctx_profile_report(self.ctx, None, 0, ctypes.byref(need)), "rb_ctx_profile_report")
        buf = ctypes.create_string_buffer(need.value)
        check(self.L.rb_ctx_profile_report(self.ctx, buf, need.value, None), "rb_ctx_profile_report")
        return json.loads(buf.value.decode())

    def ac17_msp_from_policy(self, policy):
        """policy: rabe_b200.policy.Policy -> device-resident folded MSP handle."""
        p = ctypes.c_void_p()
        check(self.L.rb_ac17_msp_from_policy(self.ctx, policy.ptr, ctypes.byref(p)), "rb_ac17_msp_from_policy")
        h = _Handle(p, self.L.rb_msp_free, self)
        n1 = ctypes.c_uint32()
        check(self.L.rb_policy_msp(policy.ptr, ctypes.byref(n1), None, None, 0, None, 0, None), "rb_policy_msp")
        h.n1 = n1.value
        return h

    def _out(self, like, nbytes):
        if _is_cuda_tensor(like):
            return torch.empty(nbytes, dtype=torch.uint8, device=like.device)
        return np.empty(nbytes, dtype=np.uint8)

    def _call(self, name, *args):
        keep, cargs, dev = [], [], []
        for a in args:
            if isinstance(a, (int, ctypes.c_void_p)) or a is None:
                cargs.append(a)
            elif isinstance(a, _Handle):
                cargs.append(a.ptr)
            else:
                ptr, k, on_dev = _as_buf(a)
                keep.append(k)
                if on_dev:
                    dev.append(k)
                cargs.append(ctypes.c_void_p(ptr))
        # Device-resident buffers: the call only enqueues on the context's stream.  Order that stream with torch's
        # current one in both directions (inputs written by torch are complete before the kernels read them; torch
        # ops issued after the call see the outputs) and tell the caching allocator that the tensors are in use on
        # the context's stream, so a temporary freed right after the call is not recycled under a running kernel.
        ext = cur = None
        if dev:
            cur = torch.cuda.current_stream(self.device)
            ext = self._ctx_stream()
            if ext.cuda_stream != cur.cuda_stream:
                ext.wait_stream(cur)
                for t in dev:
                    t.record_stream(ext)
            else:
                ext = None
        check(getattr(self.L, name)(self.ctx, *cargs), name)
        if ext is not None:
            cur.wait_stream(ext)
        return keep

    # ------------------------------------------------------------------ L0
    def fq_mul(self, a, b):
        n = _nbytes(a) // FR
        out = self._out(a, n * FR)
        self._call("rb_fq_mul_batch", a, b, n, out)
        return out

    def fr_mul(self, a, b):
        n = _nbytes(a) // FR
        out = self._out(a, n * FR)
        self._call("rb_fr_mul_batch", a, b, n, out)
        return out

    def fq_mul_chain(self, a, b, iters):
        n = _nbytes(a) // FR
        out = self._out(a, n * FR)
        self._call("rb_fq_mul_chain", a, b, n, int(iters), out)
        return out

    def _table(self, fn, base, w):
        p = ctypes.c_void_p()
        _settle(base)
        ptr, keep, _ = _as_buf(base)
        check(getattr(self.L, fn)(self.ctx, ctypes.c_void_p(ptr), int(w), ctypes.byref(p)), fn)
        return _Handle(p, self.L.rb_table_destroy, self)

    def g1_table(self, base, window_bits=16):
        return self._table("rb_g1_table_create", base, window_bits)

    def g2_table(self, base, window_bits=8):
        return self._table("rb_g2_table_create", base, window_bits)

    def gt_table(self, base, window_bits=8):
        return self._table("rb_gt_table_create", base, window_bits)

    def g1_mul_fixed(self, table, k):
        n = _nbytes(k) // FR
        out = self._out(k, n * G1)
        self._call("rb_g1_mul_fixed_batch", table, k, n, out)
        return out

    def g2_mul_fixed(self, table, k):
        n = _nbytes(k) // FR
        out = self._out(k, n * G2)
        self._call("rb_g2_mul_fixed_batch", table, k, n, out)
        return out

    def gt_pow_fixed(self, table, k):
        n = _nbytes(k) // FR
        out = self._out(k, n * GT)
        self._call("rb_gt_pow_fixed_batch", table, k, n, out)
        return out

    def g1_mul_var(self, p, k):
        n = _nbytes(k) // FR
        out = self._out(p, n * G1)
        self._call("rb_g1_mul_var_batch", p, k, n, out)
        return out

    def g2_mul_var(self, p, k):
        n = _nbytes(k) // FR
        out = self._out(p, n * G2)
        self._call("rb_g2_mul_var_batch", p, k, n, out)
        return out

    def gt_pow_var(self, a, k):
        n = _nbytes(k) // FR
        out = self._out(a, n * GT)
        self._call("rb_gt_pow_var_batch", a, k, n, out)
        return out

    def gt_mul(self, a, b):
        n = _nbytes(a) // GT
        out = self._out(a, n * GT)
        self._call("rb_gt_mul_batch", a, b, n, out)
        return out

    def gt_inverse(self, a):
        n = _nbytes(a) // GT
        out = self._out(a, n * GT)
        self._call("rb_gt_inverse_batch", a, n, out)
        return out

    _FR_OPS = {"add": 0, "sub": 1, "mul": 2, "inverse": 3, "neg": 4}

    def fr_op(self, op, a, b=None):
        """Element-wise Fr operator; `b` may be a single 32-byte scalar (broadcast)."""
        n = _nbytes(a) // FR
        out = self._out(a, n * FR)
        scalar = 1 if (b is not None and _nbytes(b) == FR and n != 1) else 0
        self._call("rb_fr_op_batch", self._FR_OPS[op], a, b, scalar, n, out)
        return out

    def g1_add(self, a, b):
        n = _nbytes(a) // G1
        out = self._out(a, n * G1)
        self._call("rb_g1_add_batch", a, b, 1 if (_nbytes(b) == G1 and n != 1) else 0, n, out)
        return out

    def g2_add(self, a, b):
        n = _nbytes(a) // G2
        out = self._out(a, n * G2)
        self._call("rb_g2_add_batch", a, b, 1 if (_nbytes(b) == G2 and n != 1) else 0, n, out)
        return out

    def share_plan(self, policy):
        """Device-resident flattening of gen_shares_policy for `policy` (rabe_b200.policy.Policy)."""
        p = ctypes.c_void_p()
        check(self.L.rb_share_plan_create(self.ctx, policy.ptr, ctypes.byref(p)), "rb_share_plan_create")
        h = _Handle(p, self.L.rb_share_plan_free, self)
        nl, nc = ctypes.c_uint32(), ctypes.c_uint32()
        check(self.L.rb_share_plan_dims(p, ctypes.byref(nl), ctypes.byref(nc)), "rb_share_plan_dims")
        h.n_leaves, h.n_coefs = nl.value, nc.value
        return h

    def shares(self, plan, secret, coeffs):
        B = _nbytes(secret) // FR
        out = self._out(secret, B * plan.n_leaves * FR)
        self._call("rb_shares_batch", plan, secret, coeffs if plan.n_coefs else None, B, out)
        return out

    def policy_coefficients(self, policy, n_leaves):
        out = np.empty(n_leaves * FR, dtype=np.uint8)
        check(self.L.rb_policy_coefficients(self.ctx, policy.ptr, ctypes.c_void_p(out.ctypes.data)), "rb_policy_coefficients")
        return out

    def g1_sum_gather(self, points, idx, offs):
        idx = np.ascontiguousarray(idx, dtype=np.uint32) if not _is_cuda_tensor(idx) else idx
        offs = np.ascontiguousarray(offs, dtype=np.uint32) if not _is_cuda_tensor(offs) else offs
        n_out = (_nbytes(offs) // 4) - 1
        out = self._out(points, n_out * G1)
        self._call("rb_g1_sum_gather_batch", points, _nbytes(points) // G1, idx, offs, n_out, out)
        return out

    def pairing_product(self, P, Q, offs):
        offs = np.ascontiguousarray(offs, dtype=np.uint32) if not _is_cuda_tensor(offs) else offs
        n = (_nbytes(offs) // 4) - 1
        out = self._out(P, n * GT)
        self._call("rb_pairing_product_batch", P, Q, offs, n, out)
        return out

    def pairing(self, P, Q):
        n = _nbytes(P) // G1
        return self.pairing_product(P, Q, np.arange(n + 1, dtype=np.uint32))

    def sha3_fr(self, strings):
        """SHA3-256 -> Fr of a list of byte strings on the device (hash/mod.rs:23-31)."""
        blobs = [bytes(x) for x in strings]
        offs = np.zeros(len(blobs) + 1, dtype=np.uint32)
        offs[1:] = np.cumsum([len(b) for b in blobs])
        data = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8)
        out = np.empty(len(blobs) * FR, dtype=np.uint8)
        self._call("rb_sha3_fr_batch", data, offs, len(blobs), out)
        return out

    # ------------------------------------------------------------------ fused BSW / LSW / AW11
    def bsw_pk_load(self, g1, g2, h, e_gg_alpha):
        p = ctypes.c_void_p()
        bufs = [_as_buf(np.frombuffer(bytes(x), dtype=np.uint8)) for x in (g1, g2, h, e_gg_alpha)]
        check(self.L.rb_bsw_pk_load(self.ctx, *[ctypes.c_void_p(b[0]) for b in bufs], ctypes.byref(p)), "rb_bsw_pk_load")
        return _Handle(p, self.L.rb_bsw_pk_free, self)

    def bsw_encrypt(self, pk, plan, leaf_hash, secret, coeffs, msg):
        B, n = _nbytes(secret) // FR, plan.n_leaves
        c, c_p = self._out(secret, B * G1), self._out(secret, B * GT)
        cy1, cy2 = self._out(secret, B * n * G1), self._out(secret, B * n * G2)
        self._call("rb_bsw_encrypt_batch", pk, plan, leaf_hash, secret, coeffs if plan.n_coefs else None, msg, B, c, c_p, cy1, cy2)
        return c, c_p, cy1, cy2

    def bsw_keygen(self, pk, beta, g2_alpha, attr_hash, r, r_j):
        B, n = _nbytes(r) // FR, _nbytes(attr_hash) // FR
        d, d1, d2 = self._out(r, B * G2), self._out(r, B * n * G1), self._out(r, B * n * G2)
        self._call("rb_bsw_keygen_batch", pk, beta, g2_alpha, attr_hash, n, r, r_j, B, d, d1, d2)
        return d, d1, d2

    def bsw_delegate(self, pk, f, d, dj_g1, dj_g2, attr_hash, r, r_j):
        B, n = _nbytes(r) // FR, _nbytes(attr_hash) // FR
        od, o1, o2 = self._out(r, B * G2), self._out(r, B * n * G1), self._out(r, B * n * G2)
        self._call("rb_bsw_delegate_batch", pk, f, d, dj_g1, dj_g2, attr_hash, n, r, r_j, B, od, o1, o2)
        return od, o1, o2

    def bsw_decrypt(self, d, dj_g1, dj_g2, c, c_p, cy_g1, cy_g2, ct_idx, sk_idx, coeff):
        B, n_k = _nbytes(c) // G1, _nbytes(dj_g1) // G1
        n = _nbytes(cy_g1) // G1 // B
        ct_idx, sk_idx = np.ascontiguousarray(ct_idx, dtype=np.uint32), np.ascontiguousarray(sk_idx, dtype=np.uint32)
        nI = len(ct_idx)
        out = self._out(c, B * GT)
        self._call("rb_bsw_decrypt_batch", d, dj_g1, dj_g2, n_k, c, c_p, cy_g1, cy_g2, n, ct_idx if nI else None, sk_idx if nI else None,
                   coeff if nI else None, nI, B, out)
        return out

    def lsw_keygen(self, g1_tab, g2_tab, plan, leaf_hash, alpha1, alpha2, coeffs, rnd):
        n = plan.n_leaves
        B = _nbytes(rnd) // FR // n
        d1, d2 = self._out(rnd, B * n * G1), self._out(rnd, B * n * G2)
        self._call("rb_lsw_keygen_batch", g1_tab, g2_tab, plan, leaf_hash, alpha1, alpha2, coeffs if plan.n_coefs else None, rnd, B, d1, d2)
        return d1, d2

    def lsw_decrypt(self, sk_d1, sk_d2, e1, e2, ej1, ct_idx, sk_idx, coeff):
        B, n_k = _nbytes(e1) // GT, _nbytes(sk_d1) // G1
        n = _nbytes(ej1) // G1 // B
        ct_idx, sk_idx = np.ascontiguousarray(ct_idx, dtype=np.uint32), np.ascontiguousarray(sk_idx, dtype=np.uint32)
        nI = len(ct_idx)
        out = self._out(e1, B * GT)
        self._call("rb_lsw_decrypt_batch", sk_d1, sk_d2, n_k, e1, e2, ej1, n, ct_idx if nI else None, sk_idx if nI else None,
                   coeff if nI else None, nI, B, out)
        return out

    def lsw_pk_load(self, g1, g2, g1_b, g1_b2, h_b, e_gg_alpha):
        p = ctypes.c_void_p()
        bufs = [_as_buf(np.frombuffer(bytes(x), dtype=np.uint8)) for x in (g1, g2, g1_b, g1_b2, h_b, e_gg_alpha)]
        check(self.L.rb_lsw_pk_load(self.ctx, *[ctypes.c_void_p(b[0]) for b in bufs], ctypes.byref(p)), "rb_lsw_pk_load")
        return _Handle(p, self.L.rb_lsw_pk_free, self)

    def lsw_encrypt(self, pk, attr_hash, secret, draws, msg):
        """rb_lsw_encrypt_batch: secret [B], draws [B][n], msg [B] -> e1 [B], e2 [B], ej1/ej2/ej3 [B][n]"""
        B, n = _nbytes(secret) // FR, _nbytes(attr_hash) // FR
        e1, e2 = self._out(secret, B * GT), self._out(secret, B * G2)
        j1, j2, j3 = (self._out(secret, B * n * G1) for _ in range(3))
        self._call("rb_lsw_encrypt_batch", pk, attr_hash, n, secret, draws, msg, B, e1, e2, j1, j2, j3)
        return e1, e2, j1, j2, j3

    def ghw11_transform(self, k_z, l_z, kx, c1, ci, di, ct_idx, sk_idx, coeff):
        B, n_k = _nbytes(c1) // G1, _nbytes(kx) // G2
        n = _nbytes(ci) // G1 // B
        ct_idx, sk_idx = np.ascontiguousarray(ct_idx, dtype=np.uint32), np.ascontiguousarray(sk_idx, dtype=np.uint32)
        nI = len(ct_idx)
        out = self._out(c1, B * GT)
        self._call("rb_ghw11_transform_batch", k_z, l_z, kx, n_k, c1, ci, di, n, ct_idx if nI else None, sk_idx if nI else None,
                   coeff if nI else None, nI, B, out)
        return out

    def ghw11_decrypt_out(self, c, t, z):
        B = _nbytes(c) // GT
        out = self._out(c, B * GT)
        self._call("rb_ghw11_decrypt_out_batch", c, t, z, B, out)
        return out

    def kem_encrypt(self, gt, nonces, payloads):
        """rabe's encrypt_symmetric for a batch on the device: key = SHA3-256(Gt bytes), AES-256-GCM.  payloads: list of
        bytes; nonces: [B][12]; returns the list of nonce | ciphertext | tag blobs."""
        B = len(payloads)
        offs = np.zeros(B + 1, dtype=np.uint32); offs[1:] = np.cumsum([len(p) for p in payloads])
        data = np.frombuffer(b"".join(payloads) or b"\0", dtype=np.uint8)
        out = np.empty(int(offs[B]) + 28 * B, dtype=np.uint8)
        self._call("rb_kem_encrypt_batch", gt, nonces, data, offs, B, out)
        raw = out.tobytes()
        return [raw[int(offs[b]) + 28 * b:int(offs[b + 1]) + 28 * (b + 1)] for b in range(B)]

    def kem_decrypt(self, gt, blobs):
        """rabe's decrypt_symmetric for a batch: returns [plaintext or None (tag mismatch)]."""
        B = len(blobs)
        offs = np.zeros(B + 1, dtype=np.uint32); offs[1:] = np.cumsum([len(p) for p in blobs])
        data = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8)
        out = np.zeros(max(1, int(offs[B])), dtype=np.uint8)
        ok = np.zeros(B, dtype=np.int32)
        self._call("rb_kem_decrypt_batch", gt, data, offs, B, out, ok.view(np.uint8))
        raw = out.tobytes()
        return [raw[int(offs[b]):int(offs[b + 1]) - 28] if ok[b] else None for b in range(B)]

    def aw11_encrypt(self, g2_tab, egg_tab, plan, pk_gt, pk_g2, s, s_coeffs, w_coeffs, r_x, msg):
        B, n = _nbytes(s) // FR, plan.n_leaves
        c0, c1 = self._out(s, B * GT), self._out(s, B * n * GT)
        c2, c3 = self._out(s, B * n * G2), self._out(s, B * n * G2)
        self._call("rb_aw11_encrypt_batch", g2_tab, egg_tab, plan, pk_gt, pk_g2, s, s_coeffs if plan.n_coefs else None,
                   w_coeffs if plan.n_coefs else None, r_x, msg, B, c0, c1, c2, c3)
        return c0, c1, c2, c3

    def aw11_decrypt(self, h, sk_k, c_0, c1, c2, c3, ct_idx, sk_idx, coeff):
        B, n_k = _nbytes(c_0) // GT, _nbytes(sk_k) // G1
        n = _nbytes(c1) // GT // B
        ct_idx, sk_idx = np.ascontiguousarray(ct_idx, dtype=np.uint32), np.ascontiguousarray(sk_idx, dtype=np.uint32)
        nI = len(ct_idx)
        out = self._out(c_0, B * GT)
        self._call("rb_aw11_decrypt_batch", h, sk_k, n_k, c_0, c1, c2, c3, n, ct_idx if nI else None, sk_idx if nI else None,
                   coeff if nI else None, nI, B, out)
        return out

    def aw11_pk_load(self, pk_gt, pk_g2):
        n = _nbytes(pk_gt) // GT
        p = ctypes.c_void_p()
        _settle(pk_gt, pk_g2)
        b1, b2 = _as_buf(pk_gt), _as_buf(pk_g2)
        check(self.L.rb_aw11_pk_load(self.ctx, ctypes.c_void_p(b1[0]), ctypes.c_void_p(b2[0]), n, ctypes.byref(p)), "rb_aw11_pk_load")
        h = _Handle(p, self.L.rb_aw11_pk_free, self)
        h.n = n
        return h

    def aw11_encrypt_pk(self, g2_tab, egg_tab, plan, pk, leaf_attr, s, s_coeffs, w_coeffs, r_x, msg):
        B, n = _nbytes(s) // FR, plan.n_leaves
        c0, c1 = self._out(s, B * GT), self._out(s, B * n * GT)
        c2, c3 = self._out(s, B * n * G2), self._out(s, B * n * G2)
        la = None if leaf_attr is None else np.ascontiguousarray(leaf_attr, dtype=np.uint32)
        self._call("rb_aw11_encrypt_pk_batch", g2_tab, egg_tab, plan, pk, la, s, s_coeffs if plan.n_coefs else None,
                   w_coeffs if plan.n_coefs else None, r_x, msg, B, c0, c1, c2, c3)
        return c0, c1, c2, c3

    # ------------------------------------------------------------------ AC17
    def ac17_setup(self, rnd):
        pk, msk = np.empty(1216, np.uint8), np.empty(512, np.uint8)
        self._call("rb_ac17_setup", rnd, pk, msk)
        return pk.tobytes(), msk.tobytes()

    def ac17_pk_load(self, pk, g1_window=16, g2_window=8, gt_window=8):
        p = ctypes.c_void_p()
        ptr, keep, _ = _as_buf(pk)
        check(self.L.rb_ac17_pk_load_ex(self.ctx, ctypes.c_void_p(ptr), int(g1_window), int(g2_window), int(gt_window), ctypes.byref(p)),
              "rb_ac17_pk_load_ex")
        return _Handle(p, self.L.rb_ac17_pk_free, self)

    def ac17_msk_load(self, msk):
        p = ctypes.c_void_p()
        ptr, keep, _ = _as_buf(msk)
        check(self.L.rb_ac17_msk_load(self.ctx, ctypes.c_void_p(ptr), ctypes.byref(p)), "rb_ac17_msk_load")
        return _Handle(p, self.L.rb_ac17_msk_free, self)

    def msp_load_batch(self, m, h_row, h_col):
        """m: int8 [n_pol][n1][n2]; one folded policy per batch item (policy_mode = distinct)."""
        m = np.ascontiguousarray(m, dtype=np.int8)
        n_pol, n1, n2 = m.shape
        p = ctypes.c_void_p()
        _settle(h_row, h_col)
        pm, k1, _ = _as_buf(m.view(np.uint8).reshape(-1))
        pr, k2, _ = _as_buf(h_row)
        pc, k3, _ = _as_buf(h_col)
        check(self.L.rb_msp_load_batch(self.ctx, n1, n2, ctypes.c_void_p(pm), ctypes.c_void_p(pr), ctypes.c_void_p(pc), n_pol, ctypes.byref(p)),
              "rb_msp_load_batch")
        h = _Handle(p, self.L.rb_msp_free, self)
        h.n1, h.n2 = n1, n2
        return h

    def msp_reload_batch(self, msp, m, h_row, h_col, h_col_shared=False):
        """Refold `msp` in place (same n_pol / n1 / n2) from new matrices and hashes; host or device buffers.
        h_col_shared: h_col is one [n2][3][2] table for every policy (the column labels are policy-independent)."""
        if not _is_cuda_tensor(m):
            m = np.ascontiguousarray(m, dtype=np.int8).view(np.uint8).reshape(-1)
        self._call("rb_msp_reload_batch", msp, m, h_row, h_col, 1 if h_col_shared else 0)
        return msp

    def sha3_fr_packed(self, data, offs, n, out=None):
        """SHA3-256 -> Fr of n strings already packed as data + uint32 offs[n+1] (host arrays or CUDA
        tensors); the data length is taken from the buffer, so device-resident inputs stay asynchronous."""
        if out is None:
            out = self._out(data, n * FR)
        self._call("rb_sha3_fr_batch_len", data, _nbytes(data), offs, int(n), out)
        return out

    def msp_load(self, m, h_row, h_col):
        m = np.ascontiguousarray(m, dtype=np.int8)
        n1, n2 = m.shape
        p = ctypes.c_void_p()
        _settle(h_row, h_col)
        pm, k1, _ = _as_buf(m.view(np.uint8))
        pr, k2, _ = _as_buf(h_row)
        pc, k3, _ = _as_buf(h_col)
        check(self.L.rb_msp_load(self.ctx, n1, n2, ctypes.c_void_p(pm), ctypes.c_void_p(pr), ctypes.c_void_p(pc), ctypes.byref(p)),
              "rb_msp_load")
        h = _Handle(p, self.L.rb_msp_free, self)
        h.n1, h.n2 = n1, n2
        return h

    def ac17_cp_encrypt(self, pk, msp, s, msg, out=None):
        B = _nbytes(s) // (2 * FR)
        if out is None:
            out = (self._out(s, B * 3 * G2), self._out(s, B * msp.n1 * 3 * G1), self._out(s, B * GT))
        self._call("rb_ac17_cp_encrypt_batch", pk, msp, s, msg, B, out[0], out[1], out[2])
        return out

    def ac17_cp_keygen(self, msk, h_attr, h_01, rnd, n):
        B = _nbytes(rnd) // ((n + 3) * FR)
        out = (self._out(rnd, B * 3 * G2), self._out(rnd, B * n * 3 * G1), self._out(rnd, B * 3 * G1))
        self._call("rb_ac17_cp_keygen_batch", msk, int(n), h_attr, h_01, rnd, B, out[0], out[1], out[2])
        return out

    def ac17_kp_keygen(self, msk, m, h_row, h_col, rnd):
        m = np.ascontiguousarray(m, dtype=np.int8)
        n1, n2 = m.shape
        B = _nbytes(rnd) // ((2 + (n2 - 1) + n1) * FR)
        out = (self._out(rnd, B * 3 * G2), self._out(rnd, B * n1 * 3 * G1))
        self._call("rb_ac17_kp_keygen_batch", msk, int(n1), int(n2), m.view(np.uint8), h_row, h_col, rnd, B, out[0], out[1])
        return out

    def ac17_sk_load(self, k_0, k, k_p):
        """Device-resident secret key with precomputed Miller lines for k_0 (fixed pairing arguments)."""
        n_k = _nbytes(k) // (3 * G1)
        p = ctypes.c_void_p()
        _settle(k_0, k, k_p)
        bufs = [_as_buf(x) for x in (k_0, k, k_p)]
        check(self.L.rb_ac17_sk_load(self.ctx, ctypes.c_void_p(bufs[0][0]), ctypes.c_void_p(bufs[1][0]), int(n_k), ctypes.c_void_p(bufs[2][0]),
                                     ctypes.byref(p)), "rb_ac17_sk_load")
        h = _Handle(p, self.L.rb_ac17_sk_free, self)
        h.n_k = n_k
        return h

    def ac17_cp_decrypt_sk(self, sk, c_0, c, c_p, n1, ct_idx, sk_idx, ct_offs=None, sk_offs=None, out=None):
        B = _nbytes(c_p) // GT
        if not _is_cuda_tensor(ct_idx):
            ct_idx = np.ascontiguousarray(ct_idx, dtype=np.uint32)
        if not _is_cuda_tensor(sk_idx):
            sk_idx = np.ascontiguousarray(sk_idx, dtype=np.uint32)
        if out is None:
            out = self._out(c_p, B * GT)
        self._call("rb_ac17_cp_decrypt_sk_batch", sk, c_0, c, int(n1), c_p, B, ct_idx, ct_offs, _nbytes(ct_idx) // 4,
                   sk_idx, sk_offs, _nbytes(sk_idx) // 4, out)
        return out

    def ac17_cp_decrypt(self, k_0, k, k_p, c_0, c, c_p, n1, ct_idx, sk_idx, ct_offs=None, sk_offs=None, out=None):
        B = _nbytes(c_p) // GT
        n_k = _nbytes(k) // (3 * G1)
        if not _is_cuda_tensor(ct_idx):
            ct_idx = np.ascontiguousarray(ct_idx, dtype=np.uint32)
        if not _is_cuda_tensor(sk_idx):
            sk_idx = np.ascontiguousarray(sk_idx, dtype=np.uint32)
        if ct_offs is not None and not _is_cuda_tensor(ct_offs):
            ct_offs = np.ascontiguousarray(ct_offs, dtype=np.uint32)
        if sk_offs is not None and not _is_cuda_tensor(sk_offs):
            sk_offs = np.ascontiguousarray(sk_offs, dtype=np.uint32)
        if out is None:
            out = self._out(c_p, B * GT)
        self._call("rb_ac17_cp_decrypt_batch", k_0, k, int(n_k), k_p, c_0, c, int(n1), c_p, B,
                   ct_idx, ct_offs, _nbytes(ct_idx) // 4, sk_idx, sk_offs, _nbytes(sk_idx) // 4, out)
        return out


def _nbytes(x):
    if _is_cuda_tensor(x):
        return x.numel() * x.element_size()
    if isinstance(x, (bytes, bytearray, memoryview)):
        return len(x)
    return np.asarray(x).nbytes
