#!/usr/bin/env python3
"""bench.py -- AC17 CP-ABE encrypt+decrypt throughput at 64 attributes (BASELINE.json metric).

One "step" = one batch of B independent cp_encrypt calls followed by cp_decrypt of those B
ciphertexts (a round trip), on synthetic inputs: a 64-attribute all-AND policy (binary AND chain,
n1 = n2 = 64 MSP rows/columns, nI = 64 pruned leaves), one key holding all 64 attributes,
seeded scalars.  Every rank works on its own batch (weak scaling, no data-path collective).

    value  round trips / s with inputs resident in HBM (CUDA events, max over ranks)
    e2e    the same through the C ABI with HOST buffers (H2D/D2H copies inside the timed region)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
`--impl reference` times the CPU oracle (the C++ restatement of rabe's operation sequence -- the
Rust reference cannot be built here) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ABE encrypt+decrypt ops/sec @64 attrs"
UNIT = "roundtrips/s"
N_ATTRS = 64
WORKLOAD = "AC17 CP-ABE, 64-attribute all-AND policy (n1=n2=64, nI=64), batch 4096 encrypt+decrypt per GPU"


def policy_text(n):
    names = ["a%d" % i for i in range(n)]
    s = '"%s"' % names[0]
    for x in names[1:]:
        s = '(%s and "%s")' % (s, x)
    return s, names


def fr_stream(seed, n):
    """n canonical Fr values from a SplitMix64 counter stream (512 bits -> mod r)."""
    import numpy as np
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    out = bytearray()
    x = (seed * 0x9E3779B97F4A7C15 + 0x1234567) & (2**64 - 1)
    for _ in range(n):
        v = 0
        for _w in range(8):
            x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
            z ^= z >> 31
            v = (v << 64) | z
        out += (v % R).to_bytes(32, "big")
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def op_model(B, n1, nI):
    """Fp-product counts per kernel launch (algorithmic work of the shipped algorithm), from the
    per-primitive counts in tests/golden/op_counts.json (counted by running the device headers on
    the host, tools/gen_op_counts.py).  Formulas are spelled out in DESIGN.md section 5."""
    c = json.load(open(os.path.join(ROOT, "tests", "golden", "op_counts.json")))
    M, nwin_g1, nwin_8 = 16, 16, 32
    rows = {
        "k_ac17_enc_rows": B * n1 * 3 * (2 + (nwin_g1 - 1) * c["g1_madd"] + 2 + c["fe_inv"] / M + 2 + 4 + 2),
        "k_ac17_enc_c0": B * 3 * ((nwin_8 - 1) * c["g2_madd"] + 3 + c["fp2_inv"] + 12 + 4),
        "k_ac17_enc_cp": B * (12 + 2 * nwin_8 * c["fp12_mul"] + 12),
        "k_g1_gather_sum": B * 3 * (nI * (2 + c["g1_on_curve"]) + (nI - 1) * c["g1_madd"] + 1 + c["fe_inv"] + 4),
        "k_ac17_dec_miller": B * 6 * (c["miller_single"] + 4 + c["g2_on_curve"]),
        "k_final_exp": B * (5 * c["fp12_mul"] + c["final_exponentiation"] + 12 + c["fp12_mul"] + 12),
    }
    return rows


def run_reference(args):
    """CPU arm: the oracle's reference-sequence AC17 (oracle/ac17.cpp) on all host cores."""
    import oracle
    from oracle import policy as opol
    from concurrent.futures import ThreadPoolExecutor
    import random
    cores = os.cpu_count() or 1
    text, names = policy_text(N_ATTRS)
    tree = opol.parse(text, opol.HUMAN)
    m, pi, _ = opol.calculate_msp(tree)
    rng = random.Random(2)
    R = oracle.pyref.R
    fr = lambda: int(rng.randrange(R)).to_bytes(32, "big")
    pk, msk = oracle.ac17_setup(b"".join(fr() for _ in range(9)))
    k0, k, kp = oracle.ac17_cp_keygen(msk, names, b"".join(fr() for _ in range(N_ATTRS + 3)))
    msg = oracle.gt_pow(oracle.pairing(oracle.g1_generator(), oracle.g2_generator()), fr())
    ok, pruned = opol.calc_pruned(names, tree)
    plist = [a for a, _ in pruned]
    sample = args.ref_sample or 2 * cores
    rnd = [fr() + fr() for _ in range(sample)]

    def one(i):
        c0, c, cp = oracle.ac17_cp_encrypt(pk, m, pi, rnd[i], msg)
        out = oracle.ac17_cp_decrypt(plist, pi, c0, c, cp, names, k0, k, kp)
        assert out == msg
        return 1

    def step():
        with ThreadPoolExecutor(max_workers=cores) as ex:      # ctypes releases the GIL
            return sum(ex.map(one, range(sample)))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += step()
    dt = time.perf_counter() - t0
    val = done / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x64-bit limbs)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm runs a bounded sample per step, all host threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d round trips per step (oracle/ac17.cpp: C++ restatement of rabe's op sequence; not the Rust binary)" % sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="round trips timed for cpu_baseline (default 2 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rabe_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from rabe_b200.engine import Engine
    from rabe_b200.policy import Policy, PolicyLanguage
    import ctypes
    from rabe_b200 import _lib

    B, n = args.batch, N_ATTRS
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = Engine(local_rank)
    eng.use_torch_stream()

    # ---- synthetic inputs through the public API (all group elements are produced by the GPU path)
    text, names = policy_text(n)
    pk, msk = eng.ac17_setup(fr_stream(2, 9))                      # same keys on every rank (seed 2)
    pkh = eng.ac17_pk_load(np.frombuffer(pk, dtype=np.uint8))
    mskh = eng.ac17_msk_load(np.frombuffer(msk, dtype=np.uint8))
    pol = Policy(text, PolicyLanguage.HumanPolicy)
    _, pi, _ = pol.msp()
    msp = eng.ac17_msp_from_policy(pol)
    h_attr, h_01 = (ctypes.c_uint8 * (192 * n))(), (ctypes.c_uint8 * 192)()
    from rabe_b200.policy import _cstrs
    _lib.check(eng.L.rb_ac17_attr_hashes(_cstrs(names), n, h_attr, h_01), "rb_ac17_attr_hashes")
    k0, k, kp = eng.ac17_cp_keygen(mskh, np.frombuffer(bytes(h_attr), dtype=np.uint8), np.frombuffer(bytes(h_01), dtype=np.uint8),
                                   fr_stream(7, n + 3), n)
    matched, nci, nsi = ctypes.c_int(), ctypes.c_uint32(), ctypes.c_uint32()
    ct_idx, sk_idx = (ctypes.c_uint32 * (2 * n))(), (ctypes.c_uint32 * (2 * n))()
    _lib.check(eng.L.rb_ac17_decrypt_lists(pol.ptr, _cstrs(names), n, _cstrs(pi), n, ctypes.byref(matched), ct_idx, 2 * n, ctypes.byref(nci),
                                           sk_idx, 2 * n, ctypes.byref(nsi)), "rb_ac17_decrypt_lists")
    assert matched.value and nci.value == n and nsi.value == n
    ct_idx_h = np.array(ct_idx[:n], dtype=np.uint32)
    sk_idx_h = np.array(sk_idx[:n], dtype=np.uint32)

    s_h = fr_stream(1000 + rank, 2 * B)                            # per-rank scalars
    gt_tab = eng.gt_table(np.frombuffer(pk[448:832], dtype=np.uint8), 8)
    msg_h = eng.gt_pow_fixed(gt_tab, fr_stream(2000 + rank, B))    # B distinct Gt "msg" values
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    s_d, msg_d = to_dev(s_h), to_dev(msg_h)
    k0_d, k_d, kp_d = to_dev(k0), to_dev(k), to_dev(kp)
    ct_idx_d, sk_idx_d = to_dev(ct_idx_h.view(np.int32)), to_dev(sk_idx_h.view(np.int32))
    c0_d = torch.empty(B * 384, dtype=torch.uint8, device=dev)
    c_d = torch.empty(B * n * 192, dtype=torch.uint8, device=dev)
    cp_d = torch.empty(B * 384, dtype=torch.uint8, device=dev)
    out_d = torch.empty(B * 384, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_resident():
        eng.ac17_cp_encrypt(pkh, msp, s_d, msg_d, out=(c0_d, c_d, cp_d))
        eng.ac17_cp_decrypt(k0_d, k_d, kp_d, c0_d, c_d, cp_d, n, ct_idx_d, sk_idx_d, out=out_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate before any timing: decrypt(encrypt(msg)) == msg for the whole batch
    step_resident()
    eng.status()
    assert bool((out_d == msg_d).all().item()), "round trip mismatch"

    for _ in range(args.warmup):
        flush.fill_(1)
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.fill_(2)                                             # L2 flush between timed iterations (not timed)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        eng.ac17_cp_encrypt(pkh, msp, s_d, msg_d, out=(c0_d, c_d, cp_d))
        e1.record(stream)
        eng.ac17_cp_decrypt(k0_d, k_d, kp_d, c0_d, c_d, cp_d, n, ct_idx_d, sk_idx_d, out=out_d)
        e2.record(stream)
        evs.append((e0, e1, e2))
    barrier()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop()
    enc_ms = sum(a.elapsed_time(b) for a, b, _ in evs)
    dec_ms = sum(b.elapsed_time(c) for _, b, c in evs)
    total_ms = enc_ms + dec_ms
    t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = [float(x) for x in t.tolist()]
    eng.status()
    value = world * B * args.steps / (total_ms / 1e3)

    # ---- per-kernel timing (CUDA events around every launch, same stream) for the roofline
    eng.profile(True)
    for _ in range(2):
        flush.fill_(3)
        step_resident()
    prof = eng.profile_report()
    eng.profile(False)
    model = op_model(B, n, n)
    per_kernel = {}
    for name, rec in prof.items():
        if name in model:
            ms = rec["ms"] / rec["launches"]
            per_kernel[name] = {"ms": ms, "fp_mul": model[name], "gfpmul_s": model[name] / ms / 1e6}
    step_kernel_ms = sum(v["ms"] for v in per_kernel.values())
    dominant = max(per_kernel, key=lambda kname: per_kernel[kname]["ms"])

    # ---- roofline denominator: the Fp-product rate of a dependent-free chain at full occupancy
    threads = 148 * 2048
    a_d = to_dev(fr_stream(5, 1024)).repeat(threads // 1024 + 1)[:32 * threads].contiguous()
    iters = 1000
    eng.fq_mul_chain(a_d, a_d, 10)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream); eng.fq_mul_chain(a_d, a_d, iters); eb.record(stream)
        torch.cuda.synchronize()
        best = min(best, ea.elapsed_time(eb))
    peak_gfpmul = threads * iters * 2 / best / 1e6

    # ---- e2e: the C ABI with HOST (pinned) buffers; copies happen inside the timed region
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    s_p, msg_p = pin(s_h), pin(msg_h)
    k0_p, k_p, kp_p = pin(k0), pin(k), pin(kp)
    c0_p = torch.empty(B * 384, dtype=torch.uint8).pin_memory()
    c_p = torch.empty(B * n * 192, dtype=torch.uint8).pin_memory()
    cp_p = torch.empty(B * 384, dtype=torch.uint8).pin_memory()
    out_p = torch.empty(B * 384, dtype=torch.uint8).pin_memory()

    def step_host():
        eng.ac17_cp_encrypt(pkh, msp, s_p.numpy(), msg_p.numpy(), out=(c0_p.numpy(), c_p.numpy(), cp_p.numpy()))
        eng.ac17_cp_decrypt(k0_p.numpy(), k_p.numpy(), kp_p.numpy(), c0_p.numpy(), c_p.numpy(), cp_p.numpy(), n, ct_idx_h, sk_idx_h,
                            out=out_p.numpy())

    step_host()
    assert bytes(out_p.numpy()) == bytes(msg_h), "e2e round trip mismatch"
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    ct_bytes = B * (384 + n * 192 + 384)
    h2d = B * 64 + B * 384 + ct_bytes + (384 + n * 192 + 192) + 8 * n     # enc inputs + dec inputs (ct, sk, lists)
    d2h = ct_bytes + B * 384

    line = None
    if rank == 0:
        hbm_bytes = B * (64 + 384) + ct_bytes + ct_bytes + B * 384         # algorithmic HBM bytes of one step
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        dom = per_kernel[dominant]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (8x32-bit limbs, Montgomery)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "policy_mode": "shared", "l2": "256 MiB flush write between timed iterations",
                       "enc_per_s": world * B * args.steps / (enc_ms / 1e3), "dec_per_s": world * B * args.steps / (dec_ms / 1e3)},
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timing": "perf_counter around synchronous C-ABI calls on pinned host buffers, max over ranks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "bound": "int-pipe (IMAD.WIDE Fp-mul rate; neither hbm nor tensor bounds this path)",
                "kernel": dominant, "achieved": dom["gfpmul_s"], "peak": peak_gfpmul, "unit": "GFpmul/s", "frac": dom["gfpmul_s"] / peak_gfpmul,
                "peak_source": "measured in this run: rb_fq_mul_chain, %d threads x %d dependent-free Montgomery products" % (threads, 2 * iters),
                "traffic": None,
                "kernel_share_of_step": dom["ms"] / step_kernel_ms,
                "step_fp_mul": sum(v["fp_mul"] for v in per_kernel.values()),
                "step_achieved": sum(v["fp_mul"] for v in per_kernel.values()) / (total_ms / args.steps) / 1e6,
                "step_frac": sum(v["fp_mul"] for v in per_kernel.values()) / (total_ms / args.steps) / 1e6 / peak_gfpmul,
                "per_kernel": per_kernel,
                "hbm": {"algorithmic_bytes_per_step": hbm_bytes, "achieved_gbs": hbm_bytes / (total_ms / args.steps) / 1e6,
                        "peak_gbs": peaks.get("hbm_gbs"), "note": "reported to show HBM is not the limiter"},
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args):
    """The oracle (kind "port": C++ restatement of rabe's op sequence) on the host cores, timed on
    a bounded sample of the same workload."""
    import oracle
    from oracle import policy as opol
    from concurrent.futures import ThreadPoolExecutor
    import random
    cores = os.cpu_count() or 1
    text, names = policy_text(N_ATTRS)
    tree = opol.parse(text, opol.HUMAN)
    m, pi, _ = opol.calculate_msp(tree)
    rng = random.Random(2)
    R = oracle.pyref.R
    fr = lambda: int(rng.randrange(R)).to_bytes(32, "big")
    pk, msk = oracle.ac17_setup(b"".join(fr() for _ in range(9)))
    k0, k, kp = oracle.ac17_cp_keygen(msk, names, b"".join(fr() for _ in range(N_ATTRS + 3)))
    msg = oracle.gt_pow(oracle.pairing(oracle.g1_generator(), oracle.g2_generator()), fr())
    ok, pruned = opol.calc_pruned(names, tree)
    plist = [a for a, _ in pruned]
    sample = args.cpu_sample or 4 * cores
    rnd = [fr() + fr() for _ in range(sample)]

    def one(i):
        c0, c, cp = oracle.ac17_cp_encrypt(pk, m, pi, rnd[i], msg)
        assert oracle.ac17_cp_decrypt(plist, pi, c0, c, cp, names, k0, k, kp) == msg
        return 1

    one(0)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        done = sum(ex.map(one, range(sample)))
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    one(0)
    single = time.perf_counter() - t1
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d round trips over %d threads (oracle/ac17.cpp reference-sequence restatement; not the Rust binary)" % (sample, cores),
            "single_thread_value": 1.0 / single}


if __name__ == "__main__":
    main()
