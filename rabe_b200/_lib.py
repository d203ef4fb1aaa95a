"""ctypes binding of librabe_b200.so (the C ABI in include/rabe_b200.h).

There is no fallback: if the shared library is missing or no CUDA device is present the engine
raises.  Nothing in this package imports the CPU oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RABE_B200_LIB", os.path.join(_HERE, "librabe_b200.so"))   # override: kernel-variant experiments

RB_OK, RB_EINVAL, RB_ENOTMEMBER, RB_EPOLICY, RB_ECUDA, RB_ENOMEM = 0, -1, -2, -3, -4, -5


class RabeB200Error(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = lib().rb_strerror(status).decode() if _LIB is not None else str(status)
        super().__init__(f"{what}: {msg} (rb_status {status})")


_LIB = None
_P = ctypes.c_void_p
_SZ = ctypes.c_size_t
_U32 = ctypes.c_uint32
_I = ctypes.c_int

_SIGS = {
    "rb_strerror": (ctypes.c_char_p, [_I]),
    "rb_version": (ctypes.c_char_p, []),
    "rb_ctx_create": (_I, [_I, ctypes.POINTER(_P)]),
    "rb_ctx_destroy": (None, [_P]),
    "rb_ctx_set_stream": (_I, [_P, _P]),
    "rb_ctx_reset_stream": (_I, [_P]),
    "rb_ctx_get_stream": (_P, [_P]),
    "rb_ctx_sync": (_I, [_P]),
    "rb_ctx_status": (_I, [_P]),
    "rb_ctx_launch_count": (ctypes.c_uint64, [_P]),
    "rb_ctx_set_g2_subgroup_check": (_I, [_P, _I]),
    "rb_ctx_set_pairing_layout": (_I, [_P, _I]),
    "rb_ctx_set_async": (_I, [_P, _I]),
    "rb_g2_check_batch": (_I, [_P, _P, _SZ]),
    "rb_lsw_pk_load": (_I, [_P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "rb_lsw_pk_free": (None, [_P]),
    "rb_lsw_encrypt_batch": (_I, [_P, _P, _P, _U32, _P, _P, _P, _SZ, _P, _P, _P, _P, _P]),
    "rb_ghw11_transform_batch": (_I, [_P, _P, _P, _P, _U32, _P, _P, _P, _U32, _P, _P, _P, _U32, _SZ, _P]),
    "rb_ghw11_decrypt_out_batch": (_I, [_P, _P, _P, _P, _SZ, _P]),
    "rb_kem_encrypt_batch": (_I, [_P, _P, _P, _P, _P, _SZ, _P]),
    "rb_kem_decrypt_batch": (_I, [_P, _P, _P, _P, _SZ, _P, _P]),
    "rb_ctx_profile": (_I, [_P, _I]),
    "rb_ctx_profile_report": (_I, [_P, _P, _SZ, ctypes.POINTER(_SZ)]),
    "rb_fq_mul_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_fr_mul_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_fq_mul_chain": (_I, [_P, _P, _P, _SZ, _I, _P]),
    "rb_g1_table_create": (_I, [_P, _P, _I, ctypes.POINTER(_P)]),
    "rb_g2_table_create": (_I, [_P, _P, _I, ctypes.POINTER(_P)]),
    "rb_gt_table_create": (_I, [_P, _P, _I, ctypes.POINTER(_P)]),
    "rb_table_destroy": (None, [_P]),
    "rb_g1_mul_fixed_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_g2_mul_fixed_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_gt_pow_fixed_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_g1_mul_var_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_g2_mul_var_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_gt_pow_var_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_g1_sum_gather_batch": (_I, [_P, _P, _SZ, _P, _P, _SZ, _P]),
    "rb_pairing_product_batch": (_I, [_P, _P, _P, _P, _SZ, _P]),
    "rb_gt_mul_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_gt_inverse_batch": (_I, [_P, _P, _SZ, _P]),
    "rb_bsw_pk_load": (_I, [_P, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "rb_bsw_pk_free": (None, [_P]),
    "rb_bsw_encrypt_batch": (_I, [_P, _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P, _P]),
    "rb_bsw_keygen_batch": (_I, [_P, _P, _P, _P, _P, _U32, _P, _P, _SZ, _P, _P, _P]),
    "rb_bsw_delegate_batch": (_I, [_P, _P, _P, _P, _P, _P, _P, _U32, _P, _P, _SZ, _P, _P, _P]),
    "rb_bsw_decrypt_batch": (_I, [_P, _P, _P, _P, _U32, _P, _P, _P, _P, _U32, _P, _P, _P, _U32, _SZ, _P]),
    "rb_lsw_keygen_batch": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _P]),
    "rb_lsw_decrypt_batch": (_I, [_P, _P, _P, _U32, _P, _P, _P, _U32, _P, _P, _P, _U32, _SZ, _P]),
    "rb_aw11_encrypt_batch": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P, _P]),
    "rb_msp_load_batch": (_I, [_P, _U32, _U32, _P, _P, _P, _SZ, ctypes.POINTER(_P)]),
    "rb_sha3_fr_batch": (_I, [_P, _P, _P, _SZ, _P]),
    "rb_sha3_fr_batch_len": (_I, [_P, _P, _SZ, _P, _SZ, _P]),
    "rb_msp_reload_batch": (_I, [_P, _P, _P, _P, _P, _I]),
    "rb_aw11_decrypt_batch": (_I, [_P, _P, _P, _U32, _P, _P, _P, _P, _U32, _P, _P, _P, _U32, _SZ, _P]),
    "rb_aw11_pk_load": (_I, [_P, _P, _P, _U32, ctypes.POINTER(_P)]),
    "rb_aw11_pk_free": (None, [_P]),
    "rb_aw11_encrypt_pk_batch": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P, _P]),
    "rb_ac17_pk_load": (_I, [_P, _P, ctypes.POINTER(_P)]),
    "rb_ac17_pk_load_ex": (_I, [_P, _P, _I, _I, _I, ctypes.POINTER(_P)]),
    "rb_ac17_pk_free": (None, [_P]),
    "rb_ac17_msk_load": (_I, [_P, _P, ctypes.POINTER(_P)]),
    "rb_ac17_msk_free": (None, [_P]),
    "rb_ac17_setup": (_I, [_P, _P, _P, _P]),
    "rb_msp_load": (_I, [_P, _U32, _U32, _P, _P, _P, ctypes.POINTER(_P)]),
    "rb_msp_free": (None, [_P]),
    "rb_ac17_cp_encrypt_batch": (_I, [_P, _P, _P, _P, _P, _SZ, _P, _P, _P]),
    "rb_ac17_cp_keygen_batch": (_I, [_P, _P, _U32, _P, _P, _P, _SZ, _P, _P, _P]),
    "rb_ac17_cp_decrypt_batch": (_I, [_P, _P, _P, _U32, _P, _P, _P, _U32, _P, _SZ, _P, _P, _SZ, _P, _P, _SZ, _P]),
    "rb_fr_op_batch": (_I, [_P, _I, _P, _P, _I, _SZ, _P]),
    "rb_g1_add_batch": (_I, [_P, _P, _P, _I, _SZ, _P]),
    "rb_g2_add_batch": (_I, [_P, _P, _P, _I, _SZ, _P]),
    "rb_share_plan_create": (_I, [_P, _P, ctypes.POINTER(_P)]),
    "rb_share_plan_free": (None, [_P]),
    "rb_share_plan_dims": (_I, [_P, ctypes.POINTER(_U32), ctypes.POINTER(_U32)]),
    "rb_shares_batch": (_I, [_P, _P, _P, _P, _SZ, _P]),
    "rb_policy_coefficients": (_I, [_P, _P, _P]),
    "rb_policy_leaf_labels": (_I, [_P, _P, _SZ, ctypes.POINTER(_SZ), ctypes.POINTER(_U32)]),
    "rb_ac17_kp_keygen_batch": (_I, [_P, _P, _U32, _U32, _P, _P, _P, _P, _SZ, _P, _P]),
    "rb_ac17_sk_load": (_I, [_P, _P, _P, _U32, _P, ctypes.POINTER(_P)]),
    "rb_ac17_sk_free": (None, [_P]),
    "rb_ac17_cp_decrypt_sk_batch": (_I, [_P, _P, _P, _P, _U32, _P, _SZ, _P, _P, _SZ, _P, _P, _SZ, _P]),
    "rb_policy_parse": (_I, [ctypes.c_char_p, _I, ctypes.POINTER(_P)]),
    "rb_policy_free": (None, [_P]),
    "rb_policy_serialize": (_I, [_P, _I, _P, _SZ, ctypes.POINTER(_SZ)]),
    "rb_policy_msp": (_I, [_P, ctypes.POINTER(_U32), ctypes.POINTER(_U32), _P, _SZ, _P, _SZ, ctypes.POINTER(_SZ)]),
    "rb_policy_satisfied": (_I, [_P, _P, _U32, ctypes.POINTER(_I)]),
    "rb_policy_prune": (_I, [_P, _P, _U32, ctypes.POINTER(_I), _P, _SZ, ctypes.POINTER(_SZ), ctypes.POINTER(_U32)]),
    "rb_hash_to_fr": (_I, [ctypes.c_char_p, _SZ, _P]),
    "rb_ac17_msp_from_policy": (_I, [_P, _P, ctypes.POINTER(_P)]),
    "rb_ac17_attr_hashes": (_I, [_P, _U32, _P, _P]),
    "rb_ac17_decrypt_lists": (_I, [_P, _P, _U32, _P, _U32, ctypes.POINTER(_I), _P, _SZ, ctypes.POINTER(_U32), _P, _SZ, ctypes.POINTER(_U32)]),
}

EXPORTS = tuple(_SIGS)

# csrc/internal.h: test hooks and A/B switches that are not part of the public header
_INTERNAL_SIGS = {
    "rb_dbg_wide_dot": (_I, [_P, _P, _P, _I, _SZ, _P]),
    "rb_dbg_w6_op": (_I, [_P, _I, _I, _P, _P, _SZ, _P]),
    "rb_dbg_fq_sqr": (_I, [_P, _P, _P, _I, _SZ, _P]),
}


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  rabe_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(_SIGS.items()) + list(_INTERNAL_SIGS.items()):
            fn = getattr(L, name)      # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status, what):
    if status != RB_OK:
        raise RabeB200Error(status, what)
