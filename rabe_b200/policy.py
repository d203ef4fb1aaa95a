"""Host-side policy helpers: thin ctypes views of the C++ policy layer in librabe_b200.so
(rabe_b200/csrc/host_policy.cpp), mirroring rabe's utils::policy / utils::secretsharing /
utils::tools helpers for the hot path (strings only)."""
import ctypes
import enum

from . import _lib
from ._lib import check
from .error import RabeError


class PolicyLanguage(enum.IntEnum):           # pest/mod.rs:17-22
    JsonPolicy = 0
    HumanPolicy = 1


def _cstrs(items):
    arr = (ctypes.c_char_p * max(len(items), 1))()
    arr[:len(items)] = [s.encode() if isinstance(s, str) else bytes(s) for s in items]
    return arr


class Policy:
    """A parsed policy tree (pest/mod.rs:40 `parse`)."""

    def __init__(self, text: str, language: PolicyLanguage):
        self.L = _lib.lib()
        self.text, self.language = text, PolicyLanguage(language)
        p = ctypes.c_void_p()
        st = self.L.rb_policy_parse(text.encode(), int(language), ctypes.byref(p))
        if st == _lib.RB_EPOLICY:
            raise RabeError("policy parse error")
        check(st, "rb_policy_parse")
        self.ptr = p

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.L.rb_policy_free(self.ptr)
                self.ptr = None
        except Exception:
            pass

    def serialize(self, language):
        need = ctypes.c_size_t()
        check(self.L.rb_policy_serialize(self.ptr, int(language), None, 0, ctypes.byref(need)), "rb_policy_serialize")
        buf = ctypes.create_string_buffer(need.value)
        check(self.L.rb_policy_serialize(self.ptr, int(language), buf, need.value, None), "rb_policy_serialize")
        return buf.value.decode()

    def msp(self):
        """AbePolicy (msp.rs:11): returns (m rows, pi, c)."""
        n1, n2, need = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_size_t()
        st = self.L.rb_policy_msp(self.ptr, ctypes.byref(n1), ctypes.byref(n2), None, 0, None, 0, ctypes.byref(need))
        if st == _lib.RB_EPOLICY:
            raise RabeError("lewko waters algorithm failed =(")
        check(st, "rb_policy_msp")
        m = (ctypes.c_int8 * (n1.value * n2.value))()
        names = ctypes.create_string_buffer(need.value)
        check(self.L.rb_policy_msp(self.ptr, None, None, m, len(m), names, need.value, None), "rb_policy_msp")
        pi = [s.decode() for s in names.raw[:need.value].split(b"\0")[:-1]]
        rows = [list(m[i * n2.value:(i + 1) * n2.value]) for i in range(n1.value)]
        return rows, pi, n2.value

    def leaf_labels(self):
        """node_index labels (secretsharing/mod.rs:74) of the leaves in DFS order."""
        need, n = ctypes.c_size_t(), ctypes.c_uint32()
        check(self.L.rb_policy_leaf_labels(self.ptr, None, 0, ctypes.byref(need), ctypes.byref(n)), "rb_policy_leaf_labels")
        buf = ctypes.create_string_buffer(max(need.value, 1))
        check(self.L.rb_policy_leaf_labels(self.ptr, buf, need.value, None, None), "rb_policy_leaf_labels")
        return [s.decode() for s in buf.raw[:need.value].split(b"\0")[:-1]]

    def satisfied(self, attrs):
        out = ctypes.c_int()
        check(self.L.rb_policy_satisfied(self.ptr, _cstrs(attrs), len(attrs), ctypes.byref(out)), "rb_policy_satisfied")
        return bool(out.value)

    def prune(self, attrs):
        """calc_pruned (secretsharing/mod.rs:143): (match, [(name, node_index)])."""
        matched, need, n = ctypes.c_int(), ctypes.c_size_t(), ctypes.c_uint32()
        arr = _cstrs(attrs)
        st = self.L.rb_policy_prune(self.ptr, arr, len(attrs), ctypes.byref(matched), None, 0, ctypes.byref(need), ctypes.byref(n))
        if st == _lib.RB_EPOLICY:
            raise RabeError("Error: Invalid policy (gate with just a single child).")
        check(st, "rb_policy_prune")
        buf = ctypes.create_string_buffer(max(need.value, 1))
        check(self.L.rb_policy_prune(self.ptr, arr, len(attrs), ctypes.byref(matched), buf, need.value, None, None), "rb_policy_prune")
        parts = [s.decode() for s in buf.raw[:need.value].split(b"\0")[:-1]]
        return bool(matched.value), list(zip(parts[0::2], parts[1::2]))


def remove_index(label: str) -> str:            # secretsharing/mod.rs:77
    return label.split("_")[0]


def sha3_hash_fr(data: str) -> bytes:         # hash/mod.rs:23
    out = (ctypes.c_uint8 * 32)()
    e = data.encode()
    check(_lib.lib().rb_hash_to_fr(e, len(e), out), "rb_hash_to_fr")
    return bytes(out)
