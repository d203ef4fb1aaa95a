#!/usr/bin/env python3
"""Throughput of the fused BSW / LSW / AW11 entry points at BASELINE.json configs 3-5 (per-GPU
share of the batch), device-resident inputs, CUDA-event timing.  Development / documentation aid:
bench.py is the contract benchmark.  Prints one JSON line per operation."""
import argparse, json, os, random, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rabe_b200.engine import Engine
from rabe_b200.policy import Policy, PolicyLanguage, remove_index, sha3_hash_fr
from rabe_b200.schemes import aw11, bsw, common, lsw

R = common.R_ORDER
u8 = lambda b: np.frombuffer(bytes(b), dtype=np.uint8)


def and_tree(groups):
    return "(" + " and ".join("(" + " and ".join('"%s"' % a for a in g) + ")" for g in groups) + ")"


def binary_tree(names, rng, p_and=0.5):
    if len(names) == 1:
        return '"%s"' % names[0]
    k = rng.randrange(1, len(names))
    return "(%s %s %s)" % (binary_tree(names[:k], rng, p_and), "and" if rng.random() < p_and else "or", binary_tree(names[k:], rng, p_and))


def frs(rng, n):
    return torch.from_numpy(np.frombuffer(b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8).copy()).cuda()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream()); fn(); b.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="scale the batch sizes (1.0 = BASELINE per-GPU batches)")
    args = ap.parse_args()
    eng = Engine(0); common.set_engine(eng)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s); eng.use_torch_stream()
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(u8(x))).cuda()
    rng = random.Random(7)
    out = []
    # ---- config 3: BSW, 128-attribute AND(16 x AND(8)), B = 4096
    B = max(1, int(4096 * args.scale))
    names = ["a%d" % i for i in range(128)]
    text = and_tree([names[8 * g:8 * g + 8] for g in range(16)])
    pk, msk = bsw.setup(common.Rng(3))
    pkh = bsw._pk_handle(pk)
    pol = Policy(text, PolicyLanguage.HumanPolicy); plan = eng.share_plan(pol); labels = pol.leaf_labels()
    leaf_hash = dev(b"".join(sha3_hash_fr(remove_index(l)) for l in labels))
    secrets, coeffs = frs(rng, B), frs(rng, B * plan.n_coefs)
    msgs = eng.gt_pow_fixed(common.TABLES.get("gt", pk.e_gg_alpha, 8), frs(rng, B))
    res = {}
    def enc(): res["ct"] = eng.bsw_encrypt(pkh, plan, leaf_hash, secrets, coeffs, msgs)
    t = timed(enc)
    out.append({"config": 3, "op": "bsw_encrypt", "attrs": 128, "batch": B, "ms": t, "ops_per_s": B / t * 1e3})
    sk = bsw.keygen(pk, msk, names, common.Rng(32))
    ok, pruned = pol.prune(names)
    z = eng.policy_coefficients(pol, len(labels)).tobytes()
    ctn = labels; skn = [x.string for x in sk.d_j]
    ct_idx = [ctn.index(j) for _, j in pruned]; sk_idx = [skn.index(k) for k, _ in pruned]
    coeff = dev(b"".join(z[32 * labels.index(j):32 * labels.index(j) + 32] for _, j in pruned))
    d, d1, d2 = dev(sk.d), dev(b"".join(x.g1 for x in sk.d_j)), dev(b"".join(x.g2 for x in sk.d_j))
    c, c_p, cy1, cy2 = res["ct"]
    def dec(): res["m"] = eng.bsw_decrypt(d, d1, d2, c, c_p, cy1, cy2, ct_idx, sk_idx, coeff)
    t = timed(dec)
    eng.status()
    assert bool((res["m"] == msgs).all().item()), "BSW round trip"
    out.append({"config": 3, "op": "bsw_decrypt", "attrs": 128, "pruned": len(ct_idx), "batch": B, "ms": t, "ops_per_s": B / t * 1e3})
    del res, c, c_p, cy1, cy2
    # ---- config 4: LSW, 256-attribute AND(16 x AND(16)), B = 16384 / 8 GPUs
    B = max(1, int(2048 * args.scale))
    names = ["a%d" % i for i in range(256)]
    text = and_tree([names[16 * g:16 * g + 16] for g in range(16)])
    pk, msk = lsw.setup(common.Rng(4))
    pol = Policy(text, PolicyLanguage.HumanPolicy); plan = eng.share_plan(pol); labels = pol.leaf_labels()
    leaf_hash = dev(b"".join(sha3_hash_fr(remove_index(l)) for l in labels))
    g1t, g2t = common.TABLES.get("g1", pk.g1, 16), common.TABLES.get("g2", pk.g2, 8)
    co, rn = frs(rng, B * plan.n_coefs), frs(rng, B * plan.n_leaves)
    res = {}
    def kg(): res["k"] = eng.lsw_keygen(g1t, g2t, plan, leaf_hash, dev(msk.alpha1), dev(msk.alpha2), co, rn)
    t = timed(kg)
    out.append({"config": 4, "op": "lsw_keygen", "attrs": 256, "batch": B, "ms": t, "ops_per_s": B / t * 1e3})
    key = lsw.keygen(pk, msk, text, PolicyLanguage.HumanPolicy, common.Rng(41))
    ct = lsw.encrypt(pk, names, b"x", common.Rng(42))
    ok, pruned = pol.prune(names)
    z = eng.policy_coefficients(pol, len(labels)).tobytes()
    skn, ctn = [x[0] for x in key.dj], [x[0] for x in ct.ej]
    ci, si = [ctn.index(nm) for nm, _ in pruned], [skn.index(nm) for nm, _ in pruned]
    coeff = dev(b"".join(z[32 * labels.index(l):32 * labels.index(l) + 32] for _, l in pruned))
    k1, k2 = dev(b"".join(x[1] for x in key.dj)), dev(b"".join(x[2] for x in key.dj))
    e1, e2, ej1 = dev(ct.e1).repeat(B), dev(ct.e2).repeat(B), dev(b"".join(x[1] for x in ct.ej)).repeat(B)
    ref = lsw.decrypt_gt(key, ct)
    def ldec(): res["m"] = eng.lsw_decrypt(k1, k2, e1, e2, ej1, ci, si, coeff)
    t = timed(ldec)
    eng.status()
    assert bytes(res["m"][:384].cpu().numpy()) == ref and bytes(res["m"][-384:].cpu().numpy()) == ref, "LSW decrypt"
    out.append({"config": 4, "op": "lsw_decrypt", "attrs": 256, "pruned": len(ci), "batch": B, "ms": t, "ops_per_s": B / t * 1e3})
    del res, e1, e2, ej1
    # ---- config 5: AW11, 8 authorities x 32 attributes, binary AND/OR tree over 256, B = 8192 / 8 GPUs
    B = max(1, int(1024 * args.scale))
    gk = aw11.setup(common.Rng(5))
    auth_names = [["AUTH%dATTR%d" % (k, j) for j in range(32)] for k in range(8)]
    auths = [aw11.authgen(gk, nm, common.Rng(50 + k)) for k, nm in enumerate(auth_names)]
    flat = [n for nm in auth_names for n in nm]
    text = binary_tree(flat, random.Random(5))
    pol = Policy(text, PolicyLanguage.HumanPolicy); plan = eng.share_plan(pol); labels = pol.leaf_labels()
    rows = [aw11.find_pk_attr([a[0] for a in auths], remove_index(l.upper())) for l in labels]
    pk_gt, pk_g2 = dev(b"".join(a[1] for a in rows)), dev(b"".join(a[2] for a in rows))
    g2t, egg = common.TABLES.get("g2", gk.g2, 8), common.TABLES.get("gt", aw11._e_gg(gk), 8)
    S, SC, WC, RX = frs(rng, B), frs(rng, B * plan.n_coefs), frs(rng, B * plan.n_coefs), frs(rng, B * plan.n_leaves)
    msgs = eng.gt_pow_fixed(egg, frs(rng, B))
    res = {}
    def aenc(): res["c"] = eng.aw11_encrypt(g2t, egg, plan, pk_gt, pk_g2, S, SC, WC, RX, msgs)
    t = timed(aenc)
    eng.status()
    out.append({"config": 5, "op": "aw11_encrypt", "rows": plan.n_leaves, "batch": B, "ms": t, "ops_per_s": B / t * 1e3})
    ref_c = [x.clone() for x in res["c"]]
    pkh = eng.aw11_pk_load(pk_gt, pk_g2)                                        # per-attribute fixed-base tables
    def aenc2(): res["c"] = eng.aw11_encrypt_pk(g2t, egg, plan, pkh, None, S, SC, WC, RX, msgs)
    t = timed(aenc2)
    eng.status()
    assert all(bool((a == b).all().item()) for a, b in zip(ref_c, res["c"])), "AW11 table path differs"
    out.append({"config": 5, "op": "aw11_encrypt_pk_tables", "rows": plan.n_leaves, "batch": B, "ms": t, "ops_per_s": B / t * 1e3})
    # aw11::decrypt of those ciphertexts with a key that holds all 256 attributes
    sk = aw11.Aw11SecretKey("alice", [])
    for (apk, amsk), names_k in zip(auths, auth_names):
        for nm in names_k:
            aw11.add_to_attribute(gk, amsk, nm, sk)
    ok, pruned = pol.prune([x[0] for x in sk.attr])
    z = eng.policy_coefficients(pol, len(labels)).tobytes()
    ct_names = [l.upper() for l in labels]
    sk_names = [x[0] for x in sk.attr]
    ci, si = [ct_names.index(l) for _, l in pruned], [sk_names.index(nm) for nm, _ in pruned]
    coeff = dev(b"".join(z[32 * labels.index(l):32 * labels.index(l) + 32] for _, l in pruned))
    hpt = eng.g1_mul_fixed(common.TABLES.get("g1", gk.g1, 16), u8(sha3_hash_fr(sk.gid)))
    skk = dev(b"".join(x[1] for x in sk.attr))
    c0, c1, c2, c3 = res["c"]
    def adec(): res["m"] = eng.aw11_decrypt(dev(hpt.tobytes()), skk, c0, c1, c2, c3, ci, si, coeff)
    t = timed(adec)
    eng.status()
    assert bool((res["m"] == msgs).all().item()), "AW11 round trip"
    out.append({"config": 5, "op": "aw11_decrypt", "rows": plan.n_leaves, "pruned": len(ci), "batch": B, "ms": t, "ops_per_s": B / t * 1e3})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
