#!/usr/bin/env python3
"""BASELINE.json configurations 3-5 (BSW @128, LSW @256, AW11 8x32) through the fused batch entry points:
device-resident inputs, CUDA-event timing, every result checked (round trip or against the per-item API mirror).

Used three ways:
  * `python tools/bench_schemes.py [--scale S]`            one JSON line per operation (development aid);
  * `bench.py` imports run_config() for the `other_configs` entries of its default line (reduced batch) and for
    `bench.py --config 3|4|5 [--gpus N]`, where the BASELINE batch (4096 / 16384 / 8192 items) is SHARDED across
    the ranks (rabe_b200.dist.shard): rank 0 draws the keys, broadcasts their bytes, every rank works on its
    contiguous slice and the fixed-size results are gathered.
Fq-product counts per item (the roofline numerators) come from tests/golden/op_counts.json like bench.py's."""
import argparse, json, os, random, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabe_b200.engine import Engine
from rabe_b200.policy import Policy, PolicyLanguage, remove_index, sha3_hash_fr
from rabe_b200.schemes import aw11, bsw, common, lsw

R = common.R_ORDER
u8 = lambda b: np.frombuffer(bytes(b), dtype=np.uint8)
BASELINE_BATCH = {3: 4096, 4: 16384, 5: 8192}
WORKLOADS = {
    3: "BSW CP-ABE, 128 attributes, root AND over 16 x AND(8) (nI = 128), batch 4096 encrypt + decrypt",
    4: "LSW KP-ABE, 256 attributes, root AND over 16 x AND(16) (nI = 256), batch 16384 keygen + decrypt",
    5: "AW11 multi-authority CP-ABE, 8 authorities x 32 attributes, seeded binary AND/OR tree over the 256, batch 8192 encrypt",
}


def and_tree(groups):
    return "(" + " and ".join("(" + " and ".join('"%s"' % a for a in g) + ")" for g in groups) + ")"


def binary_tree(names, rng, p_and=0.5):
    if len(names) == 1:
        return '"%s"' % names[0]
    k = rng.randrange(1, len(names))
    return "(%s %s %s)" % (binary_tree(names[:k], rng, p_and), "and" if rng.random() < p_and else "or", binary_tree(names[k:], rng, p_and))


def frs(rng, n, dev):
    return torch.from_numpy(np.frombuffer(b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8).copy()).to(dev)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream()); fn(); b.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def op_counts():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "op_counts.json")))


def fp_mul_model(cfg, op, n, nI):
    """Fq products per ITEM of the shipped algorithm (one-thread counts of tests/golden/op_counts.json)."""
    c = op_counts()
    g1fix, g2fix13, g2fix8, gtfix = 15 * c["g1_madd"] + c["fe_inv"] / 16 + 8, 19 * c["g2_madd"] + c["fp2_inv"] + 16, 31 * c["g2_madd"] + c["fp2_inv"] + 16, 32 * c["fp12_mul"]
    fe = c["final_exponentiation"] + c["fp12_mul"]
    if (cfg, op) == (3, "encrypt"):
        return n * (g1fix + g2fix13 + 3) + g1fix + gtfix + c["fp12_mul"]
    if (cfg, op) == (3, "decrypt"):          # nI shared-accumulator pairs + one fixed pair, one final exponentiation
        return nI * (c["miller_pair"] + 15) + c["miller_fixed"] + (nI + 1) * c["fp12_mul"] + fe
    if (cfg, op) == (4, "keygen"):
        return n * (g1fix + g2fix8 + 4)
    if (cfg, op) == (4, "decrypt"):          # four fixed pairs per accumulator + the collapsed e2 pair
        return -(-nI // 4) * c["miller_fixed4_unit"] + c["miller_single"] + (-(-nI // 4) + 1) * c["fp12_mul"] + fe      # line tables normalised to l0 = 1
    if (cfg, op) == (5, "encrypt"):          # per row: two Gt table walks + product, three G2 table walks + one addition
        return n * (2 * gtfix + c["fp12_mul"] + 3 * g2fix8 + c["g2_madd"]) + gtfix + c["fp12_mul"]
    raise KeyError((cfg, op))


class Ctx:
    def __init__(self, device=0):
        self.dev = torch.device("cuda", device)
        self.eng = Engine(device)
        common.set_engine(self.eng)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.eng.use_torch_stream()
        self.eng.set_g2_subgroup_check(False)       # keys and ciphertexts below are produced in-process (see bench.py details)

    def dev_bytes(self, x):
        return torch.from_numpy(np.ascontiguousarray(u8(x))).to(self.dev)


def draw_keys(cfg):
    """The scheme keys of configuration `cfg` (plain dataclasses of canonical bytes: picklable, so that rank 0 can draw them
    once and broadcast them to the other ranks -- bench.py --config)."""
    if cfg == 3:
        return bsw.setup(common.Rng(3))
    if cfg == 4:
        return lsw.setup(common.Rng(4))
    gk = aw11.setup(common.Rng(5))
    auth_names = [["AUTH%dATTR%d" % (k, j) for j in range(32)] for k in range(8)]
    return gk, [aw11.authgen(gk, nm, common.Rng(50 + k)) for k, nm in enumerate(auth_names)]


def run_config(cfg, ctx, B, seed=7, keys=None, reps=3):
    """Times the two operations BASELINE.json names for configuration `cfg` on B items.  keys: draw_keys(cfg) as drawn by
    rank 0 (None = draw them here).  Returns (list of result dicts, a fixed-size per-item output tensor for the gather)."""
    keys = keys if keys is not None else draw_keys(cfg)
    eng, dev = ctx.eng, ctx.dev
    rng = random.Random(seed)
    out = []
    PL = PolicyLanguage
    if cfg == 3:
        names = ["a%d" % i for i in range(128)]
        text = and_tree([names[8 * g:8 * g + 8] for g in range(16)])
        pk, msk = keys
        pkh = bsw._pk_handle(pk)
        pol = Policy(text, PL.HumanPolicy); plan = eng.share_plan(pol); labels = pol.leaf_labels()
        leaf_hash = ctx.dev_bytes(b"".join(sha3_hash_fr(remove_index(l)) for l in labels))
        secrets, coeffs = frs(rng, B, dev), frs(rng, B * plan.n_coefs, dev)
        msgs = eng.gt_pow_fixed(common.TABLES.get("gt", pk.e_gg_alpha, 8), frs(rng, B, dev))
        res = {}
        def enc(): res["ct"] = eng.bsw_encrypt(pkh, plan, leaf_hash, secrets, coeffs, msgs)
        t = timed(enc, reps)
        out.append({"config": 3, "op": "encrypt", "entry": "rb_bsw_encrypt_batch", "attrs": 128, "batch": B, "ms": t, "fp_mul_per_item": fp_mul_model(3, "encrypt", 128, 128)})
        sk = bsw.keygen(pk, msk, names, common.Rng(32))
        ok, pruned = pol.prune(names)
        z = eng.policy_coefficients(pol, len(labels)).tobytes()
        skn = [x.string for x in sk.d_j]
        ct_idx = [labels.index(j) for _, j in pruned]; sk_idx = [skn.index(k) for k, _ in pruned]
        coeff = ctx.dev_bytes(b"".join(z[32 * labels.index(j):32 * labels.index(j) + 32] for _, j in pruned))
        d, d1, d2 = ctx.dev_bytes(sk.d), ctx.dev_bytes(b"".join(x.g1 for x in sk.d_j)), ctx.dev_bytes(b"".join(x.g2 for x in sk.d_j))
        c, c_p, cy1, cy2 = res["ct"]
        def dec(): res["m"] = eng.bsw_decrypt(d, d1, d2, c, c_p, cy1, cy2, ct_idx, sk_idx, coeff)
        t = timed(dec, reps)
        eng.status()
        assert bool((res["m"] == msgs).all().item()), "BSW round trip"
        out.append({"config": 3, "op": "decrypt", "entry": "rb_bsw_decrypt_batch", "attrs": 128, "pruned": len(ct_idx), "batch": B, "ms": t,
                    "fp_mul_per_item": fp_mul_model(3, "decrypt", 128, len(ct_idx))})
        return out, res["m"]
    if cfg == 4:
        names = ["a%d" % i for i in range(256)]
        text = and_tree([names[16 * g:16 * g + 16] for g in range(16)])
        pk, msk = keys
        pol = Policy(text, PL.HumanPolicy); plan = eng.share_plan(pol); labels = pol.leaf_labels()
        leaf_hash = ctx.dev_bytes(b"".join(sha3_hash_fr(remove_index(l)) for l in labels))
        g1t, g2t = common.TABLES.get("g1", pk.g1, 16), common.TABLES.get("g2", pk.g2, 8)
        co, rn = frs(rng, B * plan.n_coefs, dev), frs(rng, B * plan.n_leaves, dev)
        a1, a2 = ctx.dev_bytes(msk.alpha1), ctx.dev_bytes(msk.alpha2)
        res = {}
        def kg(): res["k"] = eng.lsw_keygen(g1t, g2t, plan, leaf_hash, a1, a2, co, rn)
        t = timed(kg, reps)
        out.append({"config": 4, "op": "keygen", "entry": "rb_lsw_keygen_batch", "attrs": 256, "batch": B, "ms": t, "fp_mul_per_item": fp_mul_model(4, "keygen", 256, 256)})
        # B ciphertexts over the 256 attributes (one fused rb_lsw_encrypt_batch), decrypted under key 0 of the batch above
        hashes = ctx.dev_bytes(b"".join(sha3_hash_fr(a) for a in names))
        secrets, draws = frs(rng, B, dev), frs(rng, B * 256, dev)
        msgs = eng.gt_pow_fixed(common.TABLES.get("gt", pk.e_gg_alpha, 8), frs(rng, B, dev))
        e1, e2, ej1, _, _ = eng.lsw_encrypt(lsw._pk_handle(pk), hashes, secrets, draws, msgs)
        key = lsw.keygen(pk, msk, text, PL.HumanPolicy, common.Rng(41))
        ok, pruned = pol.prune(names)
        z = eng.policy_coefficients(pol, len(labels)).tobytes()
        skn = [x[0] for x in key.dj]
        ci, si = [names.index(nm) for nm, _ in pruned], [skn.index(nm) for nm, _ in pruned]
        coeff = ctx.dev_bytes(b"".join(z[32 * labels.index(l):32 * labels.index(l) + 32] for _, l in pruned))
        k1, k2 = ctx.dev_bytes(b"".join(x[1] for x in key.dj)), ctx.dev_bytes(b"".join(x[2] for x in key.dj))
        def ldec(): res["m"] = eng.lsw_decrypt(k1, k2, e1, e2, ej1, ci, si, coeff)
        t = timed(ldec, reps)
        eng.status()
        assert bool((res["m"] == msgs).all().item()), "LSW round trip"
        out.append({"config": 4, "op": "decrypt", "entry": "rb_lsw_decrypt_batch", "attrs": 256, "pruned": len(ci), "batch": B, "ms": t,
                    "fp_mul_per_item": fp_mul_model(4, "decrypt", 256, len(ci))})
        return out, res["m"]
    if cfg == 5:
        gk, auths = keys
        auth_names = [["AUTH%dATTR%d" % (k, j) for j in range(32)] for k in range(8)]
        flat = [n for nm in auth_names for n in nm]
        text = binary_tree(flat, random.Random(5))
        pol = Policy(text, PL.HumanPolicy); plan = eng.share_plan(pol); labels = pol.leaf_labels()
        rows = [aw11.find_pk_attr([a[0] for a in auths], remove_index(l.upper())) for l in labels]
        pk_gt, pk_g2 = ctx.dev_bytes(b"".join(a[1] for a in rows)), ctx.dev_bytes(b"".join(a[2] for a in rows))
        g2t, egg = common.TABLES.get("g2", gk.g2, 8), common.TABLES.get("gt", aw11._e_gg(gk), 8)
        S, SC, WC, RX = frs(rng, B, dev), frs(rng, B * plan.n_coefs, dev), frs(rng, B * plan.n_coefs, dev), frs(rng, B * plan.n_leaves, dev)
        msgs = eng.gt_pow_fixed(egg, frs(rng, B, dev))
        pkh = eng.aw11_pk_load(pk_gt, pk_g2)                                        # per-attribute fixed-base tables
        res = {}
        def aenc(): res["c"] = eng.aw11_encrypt_pk(g2t, egg, plan, pkh, None, S, SC, WC, RX, msgs)
        t = timed(aenc, reps)
        eng.status()
        out.append({"config": 5, "op": "encrypt", "entry": "rb_aw11_encrypt_pk_batch", "rows": plan.n_leaves, "batch": B, "ms": t,
                    "fp_mul_per_item": fp_mul_model(5, "encrypt", plan.n_leaves, 0)})
        # check: decrypt item 0 .. min(B, 64) with a key that holds all 256 attributes
        sk = aw11.Aw11SecretKey("alice", [])
        for (apk, amsk), names_k in zip(auths, auth_names):
            for nm in names_k:
                aw11.add_to_attribute(gk, amsk, nm, sk)
        ok, pruned = pol.prune([x[0] for x in sk.attr])
        z = eng.policy_coefficients(pol, len(labels)).tobytes()
        ct_names = [l.upper() for l in labels]; sk_names = [x[0] for x in sk.attr]
        ci, si = [ct_names.index(l) for _, l in pruned], [sk_names.index(nm) for nm, _ in pruned]
        coeff = ctx.dev_bytes(b"".join(z[32 * labels.index(l):32 * labels.index(l) + 32] for _, l in pruned))
        hpt = eng.g1_mul_fixed(common.TABLES.get("g1", gk.g1, 16), u8(sha3_hash_fr(sk.gid)))
        skk = ctx.dev_bytes(b"".join(x[1] for x in sk.attr))
        nchk = min(B, 64); nl = plan.n_leaves
        c0, c1, c2, c3 = res["c"]
        m = eng.aw11_decrypt(ctx.dev_bytes(hpt.tobytes()), skk, c0[:384 * nchk], c1[:384 * nl * nchk], c2[:128 * nl * nchk], c3[:128 * nl * nchk], ci, si, coeff)
        eng.status()
        assert bool((m == msgs[:384 * nchk]).all().item()), "AW11 round trip"
        return out, c0
    raise ValueError(cfg)


def cpu_port_sample(cfg):
    """The oracle's reference-sequence restatement (oracle/schemes.py over liboracle.so) on ONE item of the configuration,
    single thread: seconds per operation.  Checker-side code: only bench.py's cpu legs call this."""
    import oracle  # noqa: F401
    from oracle import policy as OP, schemes as OS
    rng = random.Random(99)
    draws = lambda n: iter([rng.randrange(R) for _ in range(n)])
    msg = OS.gt_random(rng.randrange(R))
    t = {}
    if cfg == 3:
        names = ["a%d" % i for i in range(128)]
        text = and_tree([names[8 * g:8 * g + 8] for g in range(16)])
        pk, msk = OS.bsw_setup(draws(8))
        t0 = time.perf_counter(); ct = OS.bsw_encrypt(pk, text, OP.HUMAN, msg, draws(400)); t["encrypt"] = time.perf_counter() - t0
        sk = OS.bsw_keygen(pk, msk, names, draws(200))
        t0 = time.perf_counter(); assert OS.bsw_decrypt(sk, ct) == msg; t["decrypt"] = time.perf_counter() - t0
    elif cfg == 4:
        names = ["a%d" % i for i in range(256)]
        text = and_tree([names[16 * g:16 * g + 16] for g in range(16)])
        pk, msk = OS.lsw_setup(draws(16))
        t0 = time.perf_counter(); sk = OS.lsw_keygen(pk, msk, text, OP.HUMAN, draws(800)); t["keygen"] = time.perf_counter() - t0
        ct = OS.lsw_encrypt(pk, names, msg, draws(400))
        t0 = time.perf_counter(); assert OS.lsw_decrypt(sk, ct) == msg; t["decrypt"] = time.perf_counter() - t0
    else:
        gk = OS.aw11_setup(draws(4))
        auth_names = [["AUTH%dATTR%d" % (k, j) for j in range(32)] for k in range(8)]
        pks = [OS.aw11_authgen(gk, nm, draws(80))[0] for nm in auth_names]
        text = binary_tree([n for nm in auth_names for n in nm], random.Random(5))
        t0 = time.perf_counter(); OS.aw11_encrypt(gk, pks, text, OP.HUMAN, msg, draws(1200)); t["encrypt"] = time.perf_counter() - t0
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="scale the per-GPU batch sizes (1.0 = 4096 / 2048 / 1024)")
    ap.add_argument("--profile", action="store_true", help="per-kernel CUDA-event times of every configuration (rb_ctx_profile; slows the run)")
    args = ap.parse_args()
    ctx = Ctx(0)
    for cfg, b in ((3, 4096), (4, 2048), (5, 1024)):
        if args.profile:
            ctx.eng.profile(True)
        res, _ = run_config(cfg, ctx, max(1, int(b * args.scale)))
        for o in res:
            o["ops_per_s"] = o["batch"] / o["ms"] * 1e3
            print(json.dumps(o))
        if args.profile:
            rep = ctx.eng.profile_report(); ctx.eng.profile(False)
            print(json.dumps({"config": cfg, "kernels": {k: {"launches": v["launches"], "ms_per_launch": round(v["ms"] / v["launches"], 4), "ms_total": round(v["ms"], 3)}
                                                           for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}}))


if __name__ == "__main__":
    main()
