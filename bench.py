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

try:
    METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]     # the reference's headline metric, verbatim
except Exception:
    METRIC = "ABE encrypt+decrypt ops/sec @64 attrs, 1/2/4/8 B200 vs ref CPU"
UNIT = "roundtrips/s"
N_ATTRS = 64
NCU_TRAFFIC_FILE = os.path.join("profiles", "r2_ncu_traffic.json")   # written by tools/gpu_profile_round.sh + tools/ncu_summary.py from one `ncu --set full` capture
N_INPUT_SETS = 4             # the timed loop rotates this many scalar / message sets (table lookups differ from step to step)
PARITY_SAMPLE = 8            # items of the benchmarked batch compared with the oracle before the timed region
WORKLOAD = "AC17 CP-ABE, 64-attribute all-AND policy (n1=n2=64, nI=64), batch 4096 encrypt+decrypt per GPU"
WORKLOAD_DISTINCT = ("AC17 CP-ABE, %d seeded random binary AND/OR policies over 64 attributes, one per batch item (n1=64, n2<=64, "
                     "mean nI=%.1f); labels hashed (SHA3-256 -> Fr) and scalar tables refolded on the device inside every encrypt step")


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


def op_model(B, n1, nI, windows=(16, 8, 8)):
    """Fp-product counts per kernel launch (algorithmic work of the shipped algorithm), from the
    per-primitive counts in tests/golden/op_counts.json (counted by running the device headers on
    the host, tools/gen_op_counts.py).  Formulas are spelled out in DESIGN.md section 5."""
    c = json.load(open(os.path.join(ROOT, "tests", "golden", "op_counts.json")))
    M = 24        # outputs per thread of the G1 fixed-base kernels: the launch picks 2..24 (whole waves; 21 at B = 4096, n1 = 64) -- charged
                  # at 24, the smallest inversion share, so the count never exceeds what ran
    nwin_g1, nwin_g2, nwin_gt = -(-256 // windows[0]), -(-256 // windows[1]), -(-256 // windows[2])
    rows = {
        "k_ac17_enc_rows": B * n1 * 3 * (2 + 6 + (nwin_g1 - 2) * c["g1_madd"] + 2 + c["fe_inv"] / M + 2 + 4 + 2),     # first addition: two affine entries (6)
        "k_ac17_enc_c0": B * 3 * ((nwin_g2 - 1) * c["g2_madd"] + 3 + c["fp2_inv"] + 12 + 4),
        "k_ac17_enc_cp": B * (12 + 2 * nwin_gt * c["fp12_mul"] + 12),
        "k_g1_gather_sum": B * 3 * (nI * (2 + c["g1_on_curve"]) + (nI - 1) * c["g1_madd"] + 1 + c["fe_inv"] + 4),
        # bench.py decrypts under a LOADED key (rb_ac17_sk_load): its fixed-argument line tables are normalised to l0 = 1
        # and the cheaper line product applies (miller_pair_unit; 264 products fewer per term)
        "k_ac17_dec_miller_pair": B * 3 * (c["miller_pair_unit"] + 4 + c["g2_on_curve"]),
        "k_final_exp": B * (2 * c["fp12_mul"] + c["final_exponentiation"] + 12 + c["fp12_mul"] + 12),
    }
    # the lane-paired kernels (two threads per item, coop.cuh) do the same algorithmic work
    rows["k_ac17_dec_miller_pair_co"] = rows["k_ac17_dec_miller_pair"]
    rows["k_ac17_dec_miller_item_co"] = B * (c["miller_pair3"] - 3 * (c["miller_pair"] - c["miller_pair_unit"]) + 3 * (4 + c["g2_on_curve"]))
    rows["k_final_exp_co"] = rows["k_final_exp"] + B * c["fp12_mul"]       # + the multiplication by the initial one
    # six-lane kernels (wide.cuh): one work item per ciphertext, its three terms on one accumulator; the final
    # exponentiation then sees ONE Miller value per item.  Charged the products of the one-thread algorithm they replace.
    rows["k_ac17_dec_item_w6"] = rows["k_ac17_dec_miller_item_co"]
    rows["k_final_exp_w6"] = B * (c["final_exponentiation"] + 12 + c["fp12_mul"] + 12)
    return rows


def ncu_traffic(kernel, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this
    command (profiles/r2_ncu_traffic.json), or None when the capture does not cover this kernel / batch size."""
    try:
        t = json.load(open(os.path.join(ROOT, NCU_TRAFFIC_FILE)))
        return t["kernels"][kernel]["dram_bytes_per_launch"] if t.get("batch") == batch else None
    except Exception:
        return None


def common_config(args, workload=None):
    """The workload description both arms print (the driver compares the two `config` objects)."""
    return {"workload": workload or WORKLOAD, "n_attrs": N_ATTRS, "batch_per_gpu": args.batch, "policy_mode": args.policy_mode,
            "l2": "GPU arm: working set > L2 -- rotating 53 MB ciphertext buffers, %d rotating input sets and the HBM-resident pk.g table; "
                  "256 MiB flush write between the serial (one batch in flight) iterations.  CPU arm: not applicable" % N_INPUT_SETS}


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
            "config": common_config(args),
            "details": {"note": "CPU arm: every step is a bounded sample of the batch (%d of %d round trips), all host threads; per-item work is "
                                "identical to the GPU arm's, so round trips/s compare directly" % (sample, args.batch)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d round trips per step (oracle/ac17.cpp: C++ restatement of rabe's op sequence; not the Rust binary)" % sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def config_line(args, cfg, world, value, ms, per_op, clocks, scaling_note):
    from tools import bench_schemes as bs
    return {"metric": "ABE %s ops/sec, BASELINE.json configuration %d" % ("+".join(o["op"] for o in per_op), cfg), "value": value, "unit": "items/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u256 (8x32-bit limbs, Montgomery)", "data": "synthetic",
            "config": {"workload": bs.WORKLOADS[cfg], "total_batch": bs.BASELINE_BATCH[cfg], "sharding": scaling_note,
                       "l2": "every timed call reads/writes its whole batch of per-item keys / ciphertexts (25-164 KB per item: 50 MB - 1.3 GB per rank, above the 126 MB L2 except at 8 ranks for config 3)"},
            "per_op": per_op, "clocks": clocks}


def run_sharded_config(args, rank, world, local_rank):
    """bench.py --config 3|4|5 [--gpus N]: the BASELINE batch of the configuration split into contiguous slices, one per
    rank (rabe_b200.dist.shard); rank 0 draws the scheme keys and broadcasts their bytes (one NCCL broadcast at setup);
    every rank runs the fused entry points on its slice; the fixed-size per-item outputs are gathered.  A step = the two
    operations the configuration names over the WHOLE batch; time = max over ranks."""
    import pickle
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rabe_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from rabe_b200 import dist as rd
    from tools import bench_schemes as bs
    rd.init("nccl", dev)
    cfg = args.config
    total = bs.BASELINE_BATCH[cfg]
    lo, hi = rd.shard(total, rank, world)
    ctx = bs.Ctx(local_rank)
    keys = pickle.loads(rd.broadcast_bytes(pickle.dumps(bs.draw_keys(cfg)) if rank == 0 else b"", dev))
    sampler = ClockSampler(local_rank); sampler.start()
    rd.barrier(dev)
    res, out_t = bs.run_config(cfg, ctx, hi - lo, seed=rd.rank_seed(cfg, rank), keys=keys, reps=max(3, min(args.steps, 10)))
    rd.barrier(dev)
    clocks = sampler.stop()
    per_item = out_t.numel() // (hi - lo)
    gathered = rd.gather_fixed(out_t[:per_item * (total // world)], dev) if total % world == 0 else out_t
    per_op, step_ms = [], 0.0
    peak = None
    for o in res:
        (t_max,) = rd.reduce_max([o["ms"]], dev)
        step_ms += t_max
        per_op.append({"op": o["op"], "entry": o["entry"], "ms": t_max, "items_per_s": total / t_max * 1e3,
                       "fp_mul_per_item": o["fp_mul_per_item"], "gfpmul_s": total * o["fp_mul_per_item"] / t_max / 1e6})
    if rank == 0:
        line = config_line(args, cfg, world, total / step_ms * 1e3, step_ms, per_op, clocks,
                           "%d items -> %d per rank (contiguous slices); keys drawn on rank 0 and broadcast (%d bytes, NCCL); outputs gathered (%d bytes per item, all_gather)"
                           % (total, hi - lo, len(pickle.dumps(keys)), per_item))
        line["gathered_items"] = int(gathered.numel() // per_item)
        line["gpu_launches"] = ctx.eng.launch_count()
        print(json.dumps(line))
    rd.finalize()


def run_reference_config(args):
    """CPU arm of --config 3|4|5: the oracle's reference-sequence restatement on one item per operation, single thread
    (oracle/schemes.py drives liboracle.so from Python; its per-call overhead is negligible next to the 0.2 - 1.5 s of group
    arithmetic), scaled to items/s per core."""
    from tools import bench_schemes as bs
    cfg = args.config
    t = bs.cpu_port_sample(cfg)
    step_s = sum(t.values())
    line = {"impl": "reference", "metric": "ABE %s ops/sec, BASELINE.json configuration %d" % ("+".join(t), cfg), "value": 1.0 / step_s, "unit": "items/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u256 (4x64-bit limbs)", "data": "synthetic",
            "config": {"workload": bs.WORKLOADS[cfg], "total_batch": bs.BASELINE_BATCH[cfg], "sharding": "CPU arm: one item per operation, one thread"},
            "cpu_baseline": {"value": 1.0 / step_s, "unit": "items/s", "cores": 1, "kind": "port", "sample": "one item per operation: " + json.dumps({k: round(v, 3) for k, v in t.items()}) + " s"},
            "e2e": {"value": 1.0 / step_s, "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any page-locked buffer is
    allocated, so that the pinned staging memory of the end-to-end path is local to the GPU's PCIe root: with 8 ranks
    streaming 2 x 55 MB per 7 ms step each, buffers that all sit on one socket send half the traffic across the
    inter-socket link.  Best effort (a container may hide the CPUs); returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        before = len(os.sched_getaffinity(0))
        ncpu = os.cpu_count() or 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        usable = ideal & os.sched_getaffinity(0)
        if usable and len(usable) < before:
            os.sched_setaffinity(0, usable)
            return "rank bound to %d of %d visible CPUs (NVML ideal affinity of GPU %d)" % (len(usable), before, local_rank)
        return "no narrower GPU-local CPU set visible (%d CPUs)" % before
    except Exception as ex:          # no NVML / restricted container: run unbound
        return "unbound (%s)" % type(ex).__name__


def other_configs(local_rank, peak_gfpmul):
    """Configurations 3-5 at their per-GPU batch inside the default line (so that they are driver-visible): fused entry
    points, device-resident inputs, CUDA events, round trips checked; Fq-product model and fraction of the measured product
    rate per operation; the oracle port timed on one item beside each."""
    from tools import bench_schemes as bs
    ctx = bs.Ctx(local_rank)
    out = {}
    for cfg, b in ((3, 4096), (4, 2048), (5, 1024)):          # the per-GPU batch of BASELINE.json's configuration (config 3: one GPU; 4 and 5: an eighth)
        res, _ = bs.run_config(cfg, ctx, b, reps=2)
        cpu = bs.cpu_port_sample(cfg)
        out["config_%d" % cfg] = {"workload": bs.WORKLOADS[cfg], "batch_timed": b, "ops": [
            {"op": o["op"], "entry": o["entry"], "ms": o["ms"], "items_per_s": b / o["ms"] * 1e3, "fp_mul_per_item": o["fp_mul_per_item"],
             "gfpmul_s": b * o["fp_mul_per_item"] / o["ms"] / 1e6, "frac": b * o["fp_mul_per_item"] / o["ms"] / 1e6 / peak_gfpmul,
             "cpu_port_items_per_s_1_thread": 1.0 / cpu[o["op"]]} for o in res]}
    return out


def distinct_line(args):
    """`--policy-mode distinct` (SURVEY 8d config 2 as written: a different seeded AND/OR policy per batch item) as a run of
    its own after the default one, so that it is part of the driver-visible line: a child process (the per-item label
    hashing / table refolding pipeline is a different step function), same windows, parity sample included."""
    cmd = [sys.executable, os.path.abspath(__file__), "--policy-mode", "distinct", "--no-cpu-baseline", "--no-other-configs", "--no-table-budget",
           "--steps", str(min(args.steps, 40)), "--warmup", str(args.warmup), "--batch", str(args.batch),
           "--g1-window", str(args.g1_window), "--g2-window", str(args.g2_window), "--gt-window", str(args.gt_window)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600).stdout
        d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
        return {"workload": d["config"]["workload"], "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                "e2e": d["e2e"]["value"], "h2d_bytes_per_step": d["e2e"]["h2d_bytes_per_step"], "parity_checked_items": d["parity_checked_items"],
                "step_fp_mul": d["roofline"]["step_fp_mul"], "step_frac": d["roofline"]["step_frac"], "gpu_launches": d["gpu_launches"]}
    except Exception as ex:
        return {"unavailable": "%s: %s" % (type(ex).__name__, ex)}


def oracle_parity_sample(pk, k0, k, kp, names, text, s_h, msg_h, rho_h, ct, out, B, n, per_item_policies=None):
    """Checker only (never timed, never on the product path): PARITY_SAMPLE seeded items of the batch the bench is
    about to time -- produced with the benchmarked table windows, batch size, device buffers and loaded key -- are
    recomputed by the reference-sequence oracle (oracle/ac17.cpp) and compared byte for byte: msg, c_0, c, c_p and
    the decrypted Gt value.  Returns the number of items compared."""
    import random
    from concurrent.futures import ThreadPoolExecutor
    import oracle
    from oracle import policy as opol
    c0, c, cp = ct
    sample = sorted(set(random.Random(4242).sample(range(B), min(B, PARITY_SAMPLE - 2)) + [0, B - 1]))
    pk, k0, k, kp = bytes(pk), bytes(k0), bytes(k), bytes(kp)
    s_b, msg_b, rho_b = bytes(s_h), bytes(msg_h), bytes(rho_h)

    def one(b):
        tree = opol.parse(per_item_policies[b] if per_item_policies else text, opol.HUMAN)
        m, pi, _ = opol.calculate_msp(tree)
        ok, pruned = opol.calc_pruned(names, tree)
        m_b = oracle.gt_pow(pk[448:832], rho_b[32 * b:32 * b + 32])
        e0, e1, e2 = oracle.ac17_cp_encrypt(pk, m, pi, s_b[64 * b:64 * b + 64], m_b)
        dec = oracle.ac17_cp_decrypt([a for a, _ in pruned], pi, e0, e1, e2, names, k0, k, kp)
        assert msg_b[384 * b:384 * b + 384] == m_b, ("msg differs from the oracle", b)
        assert c0[384 * b:384 * b + 384] == e0, ("c_0 differs from the oracle", b)
        assert c[192 * n * b:192 * n * (b + 1)] == e1, ("c differs from the oracle", b)
        assert cp[384 * b:384 * b + 384] == e2, ("c_p differs from the oracle", b)
        assert out[384 * b:384 * b + 384] == dec == m_b, ("decrypted Gt differs from the oracle", b)
        return 1

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        return sum(ex.map(one, sample))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="round trips timed for cpu_baseline (default 2 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dec-streams", type=int, default=6, help="decrypt contexts/streams in the software pipeline")
    ap.add_argument("--g1-window", type=int, default=26, help="window bits of the pk.g fixed-base table, signed digits (24: 5.9 GB, 10 additions per output; 26: 21.5 GB, 9)")
    ap.add_argument("--g2-window", type=int, default=16)
    ap.add_argument("--gt-window", type=int, default=16)
    ap.add_argument("--enc-streams", type=int, default=3, help="encrypt contexts/streams in the software pipeline")
    ap.add_argument("--policy-mode", choices=["shared", "distinct"], default="shared",
                    help="shared: one all-AND policy per batch (headline); distinct: a seeded random AND/OR tree per batch item, "
                         "hashed and folded on the device inside every encrypt step (SURVEY 8d config 2)")
    ap.add_argument("--check-g2", action="store_true", help="leave the G2 subgroup test of c_0 on inside the timed region (default: waived, "
                                                             "the ciphertexts come from this process; its cost is reported in details.g2_subgroup_check)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle comparison of a sample of the benchmarked batch (development aid)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configuration: 2 = AC17 @64 (headline, weak scaling: --batch items per GPU); 3 BSW @128, 4 LSW @256, "
                         "5 AW11 8x32 = their BASELINE batch (4096 / 16384 / 8192 items) SHARDED across the ranks (strong scaling)")
    ap.add_argument("--no-table-budget", action="store_true", help="skip the extra pipelined run under a 24-bit pk.g table (details.fixed_base_windows.narrower_table)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the reduced-batch runs of configurations 3-5 in the default line")
    ap.add_argument("--diag", action="store_true", help="also time encrypt-only and decrypt-only streams (stderr; development aid)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args) if args.config == 2 else run_reference_config(args)
        return
    if args.config != 2:
        return run_sharded_config(args, rank, world, local_rank)

    # 9 contexts x (1 + 2 side) streams: more than the default 8 hardware queues, which would serialise
    # independent streams that share a queue (+2-3 % with 32; must be set before CUDA initialises)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    # the per-item-policy variant runs as a child process BEFORE this process creates its CUDA context: a second live context
    # on the device costs the child 13 % of its resident rate (measured: 450 k beside an idle parent context, 518 k alone)
    distinct_first = None
    if world == 1 and not args.no_other_configs and args.policy_mode == "shared":
        distinct_first = distinct_line(args)
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rabe_b200 has no CPU fallback")
    sched = "one host thread per rank: asynchronous host-buffer calls (rb_ctx_set_async), CUDA events between contexts"
    numa = bind_to_gpu_numa(local_rank) if world > 1 else "single rank: unbound"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from rabe_b200 import dist as rd
    rd.init("nccl", dev)

    import ctypes
    from rabe_b200 import _lib
    from rabe_b200.engine import Engine
    from rabe_b200.policy import Policy, PolicyLanguage, _cstrs

    B, n = args.batch, N_ATTRS
    # one context (= one stream + scratch arena) for encryption, two for decryption: independent
    # batches are software-pipelined (encrypt(k+1) | decrypt(k) | tail of decrypt(k-1))
    ND = args.dec_streams
    sE, sD = torch.cuda.Stream(device=dev), [torch.cuda.Stream(device=dev) for _ in range(ND)]
    engE = Engine(local_rank)
    with torch.cuda.stream(sE):
        engE.use_torch_stream()
    NE = max(1, args.enc_streams)
    sEs = [sE] + [torch.cuda.Stream(device=dev) for _ in range(NE - 1)]
    engEs = [engE] + [Engine(local_rank) for _ in range(NE - 1)]
    for e_, s_ in zip(engEs[1:], sEs[1:]):
        with torch.cuda.stream(s_):
            e_.use_torch_stream()
    engD = [Engine(local_rank) for _ in range(ND)]
    for e_, s_ in zip(engD, sD):
        with torch.cuda.stream(s_):
            e_.use_torch_stream()

    # ---- synthetic inputs through the public API (all group elements are produced by the GPU path)
    text, names = policy_text(n)
    pk, msk = engE.ac17_setup(fr_stream(2, 9))                     # same keys on every rank (seed 2)
    t_tab = time.perf_counter()
    pkh = engE.ac17_pk_load(np.frombuffer(pk, dtype=np.uint8), args.g1_window, args.g2_window, args.gt_window)
    pk_cur = [pkh]                        # the resident pipeline encrypts under pk_cur[0] (swapped for the table-budget line below)
    pk_table_build_s = time.perf_counter() - t_tab
    nwin = lambda w: -(-256 // w)
    g1_entries = ((1 << (args.g1_window - 1)) + 32) if args.g1_window > 12 else (1 << args.g1_window)     # signed digits above 12 bits
    pk_table_bytes = nwin(args.g1_window) * g1_entries * 64 + 3 * (nwin(args.g2_window) << args.g2_window) * 128 + 2 * (nwin(args.gt_window) << args.gt_window) * 384
    mskh = engE.ac17_msk_load(np.frombuffer(msk, dtype=np.uint8))
    pol = Policy(text, PolicyLanguage.HumanPolicy)
    _, pi, _ = pol.msp()
    msp = engE.ac17_msp_from_policy(pol)
    h_attr, h_01 = (ctypes.c_uint8 * (192 * n))(), (ctypes.c_uint8 * 192)()
    _lib.check(engE.L.rb_ac17_attr_hashes(_cstrs(names), n, h_attr, h_01), "rb_ac17_attr_hashes")
    k0, k, kp = engE.ac17_cp_keygen(mskh, np.frombuffer(bytes(h_attr), dtype=np.uint8), np.frombuffer(bytes(h_01), dtype=np.uint8),
                                    fr_stream(7, n + 3), n)
    matched, nci, nsi = ctypes.c_int(), ctypes.c_uint32(), ctypes.c_uint32()
    ct_idx, sk_idx = (ctypes.c_uint32 * (2 * n))(), (ctypes.c_uint32 * (2 * n))()
    _lib.check(engE.L.rb_ac17_decrypt_lists(pol.ptr, _cstrs(names), n, _cstrs(pi), n, ctypes.byref(matched), ct_idx, 2 * n, ctypes.byref(nci),
                                            sk_idx, 2 * n, ctypes.byref(nsi)), "rb_ac17_decrypt_lists")
    assert matched.value and nci.value == n and nsi.value == n
    ct_idx_h = np.array(ct_idx[:n], dtype=np.uint32)
    sk_idx_h = np.array(sk_idx[:n], dtype=np.uint32)

    DISTINCT = args.policy_mode == "distinct"
    if DISTINCT:
        # ---- 4096 seeded random binary AND/OR trees over the same 64 leaves (policy_mode=distinct).
        # Per step the device hashes every row/column label (rb_sha3_fr_batch_len), refolds the per-item scalar
        # tables in place (rb_msp_reload_batch) and encrypts item b under policy b; decrypt gathers per-item lists.
        import random as _random
        prng = _random.Random(rd.rank_seed(2000, rank))

        def rand_tree(nm):
            if len(nm) == 1:
                return '"%s"' % nm[0]
            cut = prng.randrange(1, len(nm))
            return "(%s %s %s)" % (rand_tree(nm[:cut]), "and" if prng.random() < 0.5 else "or", rand_tree(nm[cut:]))
        pols_text = []
        m_all = np.zeros((B, n, n), dtype=np.int8)                  # n2 <= n: padded with zero columns
        row_strs, ct_l, sk_l, ct_o, sk_o = [], [], [], [0], [0]
        cap = 4 * n
        ci, si = (ctypes.c_uint32 * cap)(), (ctypes.c_uint32 * cap)()
        for b in range(B):
            pols_text.append(rand_tree(names))
            pb = Policy(pols_text[-1], PolicyLanguage.HumanPolicy)
            mm, pib, cb = pb.msp()
            assert len(pib) == n and cb <= n
            m_all[b, :, :cb] = np.asarray(mm, dtype=np.int8).reshape(n, cb)
            row_strs += [("%s%d%d" % (nm, l, t)).encode() for nm in pib for l in range(3) for t in range(2)]
            _lib.check(engE.L.rb_ac17_decrypt_lists(pb.ptr, _cstrs(names), n, _cstrs(pib), n, ctypes.byref(matched), ci, cap, ctypes.byref(nci),
                                                    si, cap, ctypes.byref(nsi)), "rb_ac17_decrypt_lists")
            assert matched.value
            ct_l += ci[:nci.value]; sk_l += si[:nsi.value]
            ct_o.append(len(ct_l)); sk_o.append(len(sk_l))
        # the column labels are the same for every policy (ac17:305-328): hashed once per step, shared by the batch
        col_strs = [("0%d%d%d" % (j + 1, l, t)).encode() for j in range(n) for l in range(3) for t in range(2)]
        pack = lambda strs: (np.frombuffer(b"".join(strs), dtype=np.uint8).copy(),
                             np.concatenate([[0], np.cumsum([len(x) for x in strs])]).astype(np.uint32))
        row_h, rowo_h = pack(row_strs)
        col_h, colo_h = pack(col_strs)
        n_row, n_col = len(row_strs), len(col_strs)
        del row_strs, col_strs
        ctl_h, sko_h = np.array(ct_l, dtype=np.uint32), np.array(sk_o, dtype=np.uint32)
        skl_h, cto_h = np.array(sk_l, dtype=np.uint32), np.array(ct_o, dtype=np.uint32)
        mean_nI = len(ct_l) / B
        m_flat_h = m_all.view(np.uint8).reshape(-1)

    seed = rd.rank_seed(1000, rank)
    gt_tab = engE.gt_table(np.frombuffer(pk[448:832], dtype=np.uint8), 8)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    NS = N_INPUT_SETS
    rho_hs = [fr_stream(seed + 1 + 2 * i, B) for i in range(NS)]
    s_hs = [fr_stream(seed + 2 * i, 2 * B) for i in range(NS)]                  # per-rank scalars, NS independent sets
    msg_hs = [engE.gt_pow_fixed(gt_tab, r_) for r_ in rho_hs]                   # B distinct Gt "msg" values per set
    s_ds, msg_ds = [to_dev(x) for x in s_hs], [to_dev(x) for x in msg_hs]
    s_h, msg_h, s_d, msg_d = s_hs[0], msg_hs[0], s_ds[0], msg_ds[0]
    k0_d, k_d, kp_d = to_dev(k0), to_dev(k), to_dev(kp)
    ct_idx_d, sk_idx_d = to_dev(ct_idx_h.view(np.int32)), to_dev(sk_idx_h.view(np.int32))
    NBUF = ND + NE
    cts = [(torch.empty(B * 384, dtype=torch.uint8, device=dev), torch.empty(B * n * 192, dtype=torch.uint8, device=dev),
            torch.empty(B * 384, dtype=torch.uint8, device=dev)) for _ in range(NBUF)]
    outs = [torch.empty(B * 384, dtype=torch.uint8, device=dev) for _ in range(ND)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    torch.cuda.synchronize()

    if DISTINCT:
        i32 = lambda a: to_dev(a.view(np.int32))
        row_d, rowo_d, col_d, colo_d, m_d = to_dev(row_h), i32(rowo_h), to_dev(col_h), i32(colo_h), to_dev(m_flat_h)
        ctl_d, cto_d, skl_d, sko_d = i32(ctl_h), i32(cto_h), i32(skl_h), i32(sko_h)
        hrow_d = [torch.empty(n_row * 32, dtype=torch.uint8, device=dev) for _ in range(NE)]
        hcol_d = [torch.empty(n_col * 32, dtype=torch.uint8, device=dev) for _ in range(NE)]
        msps = []
        for e_, s_, hr, hc in zip(engEs, sEs, hrow_d, hcol_d):      # one in-place-refolded handle per encrypt context
            with torch.cuda.stream(s_):
                e_.sha3_fr_packed(row_d, rowo_d, n_row, out=hr); e_.sha3_fr_packed(col_d, colo_d, n_col, out=hc)
                msps.append(e_.msp_load_batch(m_all, hr, hc.repeat(B)))
        torch.cuda.synchronize()

    def enc(buf, e=0, iset=0):
        if DISTINCT:
            engEs[e].sha3_fr_packed(row_d, rowo_d, n_row, out=hrow_d[e])
            engEs[e].sha3_fr_packed(col_d, colo_d, n_col, out=hcol_d[e])
            engEs[e].msp_reload_batch(msps[e], m_d, hrow_d[e], hcol_d[e], h_col_shared=True)
            engEs[e].ac17_cp_encrypt(pk_cur[0], msps[e], s_ds[iset], msg_ds[iset], out=cts[buf])
            return
        engEs[e].ac17_cp_encrypt(pk_cur[0], msp, s_ds[iset], msg_ds[iset], out=cts[buf])

    skh = [e_.ac17_sk_load(k0, k, kp) for e_ in engD]             # device-resident key + fixed-argument lines, per context

    def dec(d, buf):
        if DISTINCT:
            engD[d].ac17_cp_decrypt_sk(skh[d], cts[buf][0], cts[buf][1], cts[buf][2], n, ctl_d, skl_d, ct_offs=cto_d, sk_offs=sko_d, out=outs[d])
            return
        engD[d].ac17_cp_decrypt_sk(skh[d], cts[buf][0], cts[buf][1], cts[buf][2], n, ct_idx_d, sk_idx_d, out=outs[d])

    def run_pipelined(steps):
        """K independent round trips; encrypt(k) on sEs[k % NE], decrypt(k) on sD[k % ND]; ct buffers rotate."""
        ev_dec = []
        for kk in range(steps):
            buf, d, e = kk % NBUF, kk % ND, kk % NE
            if kk >= NBUF:
                sEs[e].wait_event(ev_dec[kk - NBUF])                # ct[buf] is free again
            with torch.cuda.stream(sEs[e]):
                enc(buf, e, kk % NS)
                ev = torch.cuda.Event(); ev.record(sEs[e])
            sD[d].wait_event(ev)
            with torch.cuda.stream(sD[d]):
                dec(d, buf)
                e2 = torch.cuda.Event(); e2.record(sD[d])
            ev_dec.append(e2)
        return ev_dec

    def run_serial(steps, timed=False):
        evs = []
        with torch.cuda.stream(sE):
            for _ in range(steps):
                flush.fill_(1)                                      # L2 flush between serial iterations (not timed)
                a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record(sE); enc(0); b.record(sE)
                sD[0].wait_event(b)
                with torch.cuda.stream(sD[0]):
                    dec(0, 0)
                    c.record(sD[0])
                sE.wait_event(c)
                evs.append((a, b, c))
        return evs

    # The ciphertexts decrypted here were produced in this process a moment ago: the decrypt contexts waive the
    # G2 subgroup test of c_0 (rb_ctx_set_g2_subgroup_check; the reference's cp_decrypt takes typed, already
    # validated G2 values and does not re-check them either).  `details.g2_subgroup_check` reports the cost of leaving it on.
    for e_ in engD:
        e_.set_g2_subgroup_check(args.check_g2)

    # ---- correctness gate before any timing: decrypt(encrypt(msg)) == msg for the whole batch, and a seeded sample
    # of THIS configuration's ciphertexts (windows, batch size, device buffers, loaded key) against the oracle
    run_serial(1)
    torch.cuda.synchronize()
    engE.status(); engD[0].status()
    assert bool((outs[0] == msg_d).all().item()), "round trip mismatch"
    parity_checked = 0
    if not args.no_parity_check and rank == 0:
        parity_checked = oracle_parity_sample(pk, k0, k, kp, names, text, s_h, msg_h, rho_hs[0], [x.cpu().numpy().tobytes() for x in cts[0]],
                                              outs[0].cpu().numpy().tobytes(), B, n, per_item_policies=(pols_text if DISTINCT else None))

    def timed_pipelined(steps):
        """CUDA events around `steps` pipelined round trips (recorded on sE; every other stream waits for the start
        event and sE waits for the last decrypts), bracketed by barriers + device synchronisation."""
        rd.barrier(dev)
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record(sE)
        for s_ in sD + sEs[1:]:
            s_.wait_event(t_start)
        ev_dec = run_pipelined(steps)
        for e2 in ev_dec[-ND:]:
            sE.wait_event(e2)
        t_end.record(sE)
        rd.barrier(dev)
        return t_start.elapsed_time(t_end)

    run_pipelined(max(args.warmup, 3))
    rd.barrier(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sum(e_.launch_count() for e_ in engEs + engD)
    total_ms = timed_pipelined(args.steps)
    launches = sum(e_.launch_count() for e_ in engEs + engD) - launches0
    clocks = sampler.stop()
    for e_ in engEs + engD:
        e_.status()
    assert bool((outs[(args.steps - 1) % ND] == msg_ds[(args.steps - 1) % NS]).all().item()), "round trip mismatch after the timed region"
    value, total_ms = rd.throughput(B, args.steps, total_ms, dev)

    # the same pipeline with the G2 subgroup test of every c_0 member inside the step (what a caller pays for
    # ciphertexts of unknown origin), and back
    for e_ in engD:
        e_.set_g2_subgroup_check(not args.check_g2)
    run_pipelined(ND)
    alt_steps = max(ND, min(args.steps, 40))
    alt_ms = timed_pipelined(alt_steps)
    for e_ in engD:
        e_.status()
        e_.set_g2_subgroup_check(args.check_g2)
    alt_value, _ = rd.throughput(B, alt_steps, alt_ms, dev)

    # table budget: the same pipeline under the SAME key loaded with a narrower pk.g window (24 bits = 5.9 GB instead of
    # 21.5 GB at 26: one more mixed addition per output) -- what a deployment that keeps many keys resident would run
    budget = None
    if args.g1_window > 24 and not args.no_table_budget:
        t0b = time.perf_counter()
        pk_small = engE.ac17_pk_load(np.frombuffer(pk, dtype=np.uint8), 24, args.g2_window, args.gt_window)
        small_build_s = time.perf_counter() - t0b
        pk_cur[0] = pk_small
        run_pipelined(ND)
        small_ms = timed_pipelined(alt_steps)
        for e_ in engEs + engD:
            e_.status()
        assert bool((outs[(alt_steps - 1) % ND] == msg_ds[(alt_steps - 1) % NS]).all().item()), "round trip mismatch (24-bit table)"
        small_value, _ = rd.throughput(B, alt_steps, small_ms, dev)
        pk_cur[0] = pkh
        pk_small.close()
        budget = {"g1_bits": 24, "pk_table_bytes": 11 * ((1 << 23) + 32) * 64 + 3 * (nwin(args.g2_window) << args.g2_window) * 128 + 2 * (nwin(args.gt_window) << args.gt_window) * 384,
                  "pk_table_build_s": small_build_s, "roundtrips_per_s": small_value}

    if args.diag:
        def timed(fn):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(sE); fn(); torch.cuda.synchronize(); b.record(sE); torch.cuda.synchronize()
            return a.elapsed_time(b)

        def enc_only():
            for kk in range(args.steps):
                with torch.cuda.stream(sEs[kk % NE]):
                    enc(kk % NBUF, kk % NE, kk % NS)

        def dec_only():
            for kk in range(args.steps):
                with torch.cuda.stream(sD[kk % ND]):
                    dec(kk % ND, kk % NBUF)
        for s_ in sD + sEs[1:]:
            s_.wait_stream(sE)
        print(json.dumps({"diag": {"enc_only_ms_per_batch": timed(enc_only) / args.steps, "dec_only_ms_per_batch": timed(dec_only) / args.steps,
                                   "both_ms_per_batch": total_ms / args.steps}}), file=sys.stderr)

    # ---- serial (one batch at a time) latency, for reference
    evs = run_serial(3)
    torch.cuda.synchronize()
    enc_ms = min(a.elapsed_time(b) for a, b, _ in evs)
    dec_ms = min(b.elapsed_time(c) for _, b, c in evs)

    # ---- per-kernel timing (CUDA events around every launch, one batch in flight) for the roofline.  The timed region runs
    # nine batches at once, where the AUTO policy selects the two-lane (throughput) pairing kernels: they are the ones
    # profiled for `roofline`; the six-lane (latency) kernels AUTO picks for a lone batch are reported beside them.
    def profile_pass(layout):
        engD[0].set_pairing_layout(layout)
        engE.profile(True); engD[0].profile(True)
        evs_ = run_serial(2)
        torch.cuda.synchronize()
        prof_ = {}
        for rep in (engE.profile_report(), engD[0].profile_report()):
            for kname, rec in rep.items():
                prof_[kname] = rec
        engE.profile(False); engD[0].profile(False)
        engD[0].set_pairing_layout(engD[0].PAIRING_AUTO)
        return prof_, min(b_.elapsed_time(c_) for _, b_, c_ in evs_)
    prof_lat, dec_ms_latency = profile_pass(engD[0].PAIRING_LATENCY)
    prof, dec_ms_throughput = profile_pass(engD[0].PAIRING_THROUGHPUT)
    if args.diag:
        print(json.dumps({"diag_kernels": {k_: {"ms_per_launch": r_["ms"] / r_["launches"], "launches": r_["launches"]} for k_, r_ in prof.items()}}), file=sys.stderr)
    model = op_model(B, n, int(round(mean_nI)) if DISTINCT else n, (args.g1_window, args.g2_window, args.gt_window))
    per_kernel = {}
    for name, rec in prof.items():
        if name in model:
            ms = rec["ms"] / rec["launches"]
            per_kernel[name] = {"ms": ms, "fp_mul": model[name], "gfpmul_s": model[name] / ms / 1e6}
    step_kernel_ms = sum(v["ms"] for v in per_kernel.values())
    dominant = max(per_kernel, key=lambda kname: per_kernel[kname]["ms"])
    latency_kernels = {name: {"ms": rec["ms"] / rec["launches"], "fp_mul": model[name], "gfpmul_s": model[name] / (rec["ms"] / rec["launches"]) / 1e6}
                       for name, rec in prof_lat.items() if name in model and name not in per_kernel}

    # ---- roofline denominator: the Fp-product rate of a dependent-free chain at full occupancy
    threads = 148 * 2048
    a_d = to_dev(fr_stream(5, 1024)).repeat(threads // 1024 + 1)[:32 * threads].contiguous()
    iters = 1000
    best = 1e9
    with torch.cuda.stream(sE):
        engE.fq_mul_chain(a_d, a_d, 10)
        for _ in range(3):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(sE); engE.fq_mul_chain(a_d, a_d, iters); eb.record(sE)
            torch.cuda.synchronize()
            best = min(best, ea.elapsed_time(eb))
    peak_gfpmul = threads * iters * 2 / best / 1e6

    # ---- e2e: the C ABI with HOST (pinned) buffers; copies happen inside the timed region
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    s_ps, msg_ps = [pin(x) for x in s_hs], [pin(x) for x in msg_hs]
    s_p, msg_p = s_ps[0], msg_ps[0]
    k0_p, k_p, kp_p = pin(k0), pin(k), pin(kp)
    c0_p = torch.empty(B * 384, dtype=torch.uint8).pin_memory()
    c_p = torch.empty(B * n * 192, dtype=torch.uint8).pin_memory()
    cp_p = torch.empty(B * 384, dtype=torch.uint8).pin_memory()
    out_p = torch.empty(B * 384, dtype=torch.uint8).pin_memory()

    # Host pipeline: the C-ABI calls on host buffers are synchronous, so independent batches are
    # overlapped with one host thread per context (ctypes releases the GIL): threads E0.. encrypt
    # batch k+1 while threads D0..D{ND-1} decrypt earlier batches.  Every call copies its inputs
    # host->device and its results device->host inside the timed region.
    import queue
    NH = ND + NE
    hbufs = [(torch.empty(B * 384, dtype=torch.uint8).pin_memory(), torch.empty(B * n * 192, dtype=torch.uint8).pin_memory(),
              torch.empty(B * 384, dtype=torch.uint8).pin_memory()) for _ in range(NH)]
    houts = [torch.empty(B * 384, dtype=torch.uint8).pin_memory() for _ in range(ND)]
    del c0_p, c_p, cp_p, out_p

    if DISTINCT:          # page-locked copies of the per-step policy inputs (pageable memory would make every H2D copy a blocking staged one)
        row_p, rowo_p, col_p, colo_p, m_p = [pin(x).numpy() for x in (row_h, rowo_h, col_h, colo_h, m_flat_h)]
        ctl_p, skl_p, cto_p, sko_p = [pin(x).numpy() for x in (ctl_h, skl_h, cto_h, sko_h)]

    def enc_host(buf, e=0, iset=0):
        if DISTINCT:      # labels, offsets and matrices travel host->device inside every call
            hr = engEs[e].sha3_fr_packed(row_p, rowo_p, n_row, out=hrow_d[e])
            hc = engEs[e].sha3_fr_packed(col_p, colo_p, n_col, out=hcol_d[e])
            engEs[e].msp_reload_batch(msps[e], m_p, hr, hc, h_col_shared=True)
            engEs[e].ac17_cp_encrypt(pkh, msps[e], s_ps[iset].numpy(), msg_ps[iset].numpy(), out=tuple(x.numpy() for x in hbufs[buf]))
            return
        engEs[e].ac17_cp_encrypt(pkh, msp, s_ps[iset].numpy(), msg_ps[iset].numpy(), out=tuple(x.numpy() for x in hbufs[buf]))

    def dec_host(d, buf):
        if DISTINCT:
            engD[d].ac17_cp_decrypt_sk(skh[d], hbufs[buf][0].numpy(), hbufs[buf][1].numpy(), hbufs[buf][2].numpy(), n, ctl_p, skl_p,
                                       ct_offs=cto_p, sk_offs=sko_p, out=houts[d].numpy())
            return
        engD[d].ac17_cp_decrypt_sk(skh[d], hbufs[buf][0].numpy(), hbufs[buf][1].numpy(), hbufs[buf][2].numpy(), n, ct_idx_h, sk_idx_h,
                                   out=houts[d].numpy())

    def run_host_pipeline(steps):
        """ONE host thread drives every context: with rb_ctx_set_async the host-buffer calls only enqueue (H2D copies of the
        pinned inputs, kernels, D2H copies of the results) and return; CUDA events order the contexts -- decrypt(k) waits for
        encrypt(k)'s ciphertext to be back in the pinned host buffer, encrypt(k + NH) waits until decrypt(k) has read it."""
        ev_dec, last = [], {}
        for kk in range(steps):
            buf, d, e = kk % NH, kk % ND, kk % NE
            if kk >= NH:
                sEs[e].wait_event(ev_dec[kk - NH])
            with torch.cuda.stream(sEs[e]):
                enc_host(buf, e, kk % NS)
                ev = torch.cuda.Event(); ev.record(sEs[e])
            sD[d].wait_event(ev)
            with torch.cuda.stream(sD[d]):
                dec_host(d, buf)
                e2 = torch.cuda.Event(); e2.record(sD[d])
            ev_dec.append(e2)
            last[d] = kk % NS
        for e_ in engEs + engD:
            e_.status()                      # synchronises the context and returns its sticky status
        return last

    for e_ in engEs + engD:
        e_.set_async(True)
    last = run_host_pipeline(2 * ND)
    for d, iset in last.items():
        assert bytes(houts[d].numpy()) == bytes(msg_hs[iset]), "e2e round trip mismatch"
    rd.barrier(dev)
    t0 = time.perf_counter()
    last = run_host_pipeline(args.steps)
    rd.barrier(dev)
    e2e_s = time.perf_counter() - t0
    for d, iset in last.items():
        assert bytes(houts[d].numpy()) == bytes(msg_hs[iset]), "e2e round trip mismatch after the timed region"
    (e2e_s,) = rd.reduce_max([e2e_s], dev)
    for e_ in engEs + engD:
        e_.set_async(False)
    # one batch at a time, for reference
    t1 = time.perf_counter()
    for _ in range(3):
        enc_host(0); dec_host(0, 0)
    e2e_serial_s = (time.perf_counter() - t1) / 3
    ct_bytes = B * (384 + n * 192 + 384)
    h2d = B * 64 + B * 384 + ct_bytes + (384 + n * 192 + 192) + 8 * n     # enc inputs + dec inputs (ct, sk, lists)
    d2h = ct_bytes + B * 384
    if DISTINCT:
        h2d += row_h.nbytes + rowo_h.nbytes + col_h.nbytes + colo_h.nbytes + m_flat_h.nbytes + ctl_h.nbytes + skl_h.nbytes + cto_h.nbytes + sko_h.nbytes - 8 * n

    if rank == 0:
        hbm_bytes = B * (64 + 384) + ct_bytes + ct_bytes + B * 384         # algorithmic HBM bytes of one step
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        dom = per_kernel[dominant]
        step_mul = sum(v["fp_mul"] for v in per_kernel.values())
        ms_per_step = total_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (8x32-bit limbs, Montgomery)", "data": "synthetic",
            "config": common_config(args, None if not DISTINCT else WORKLOAD_DISTINCT % (B, mean_nI)),
            "parity_checked_items": parity_checked,
            "details": {"fixed_base_windows": {"g1_bits": args.g1_window, "g2_bits": args.g2_window, "gt_bits": args.gt_window,
                                               "pk_table_bytes": pk_table_bytes, "pk_table_build_s": pk_table_build_s,
                                               "note": "pk tables built once per key, outside the timed region",
                                               "narrower_table": budget},
                        "ciphertext_buffers": NBUF,
                        "pipeline": "independent batches overlap on %d encrypt + %d decrypt CUDA streams (one rb_ctx each); CUDA_DEVICE_MAX_CONNECTIONS=%s" % (NE, ND, os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS")),
                        "parity": "%d seeded items of the benchmarked batch (these windows, B, device buffers, loaded key) == oracle/ac17.cpp byte for byte before timing; whole batch round-trips" % parity_checked,
                        "g2_subgroup_check": {"in_timed_region": bool(args.check_g2),
                                              "why": "c_0 was produced in-process; the reference's cp_decrypt takes typed, already validated G2 values (ac17/mod.rs:385)",
                                              ("roundtrips_per_s_without_check" if args.check_g2 else "roundtrips_per_s_with_check"): alt_value},
                        "serial_enc_ms": enc_ms, "serial_dec_ms": dec_ms, "serial_roundtrips_per_s": B / ((enc_ms + dec_ms) / 1e3),
                        "pairing_layouts": {"policy": "rb_ctx_set_pairing_layout AUTO: two-lane (throughput) kernels in the pipelined timed region, six-lane (latency) kernels for a lone batch",
                                            "one_batch_decrypt_ms": {"latency": dec_ms_latency, "throughput": dec_ms_throughput},
                                            "latency_kernels": {k_: dict(v_, frac=v_["gfpmul_s"] / peak_gfpmul) for k_, v_ in latency_kernels.items()}}},
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "serial_roundtrips_per_s": B / e2e_serial_s,
                    "host_wait": sched, "cpu_binding": numa,
                    "timing": "perf_counter around the whole host pipeline (C-ABI calls on pinned host buffers, asynchronous mode, ONE host thread per rank; ends with a synchronisation of every context), max over ranks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "bound": "int-pipe (IMAD.WIDE Fp-mul rate; neither hbm nor tensor bounds this path)",
                "kernel": dominant, "achieved": dom["gfpmul_s"], "peak": peak_gfpmul, "unit": "GFpmul/s", "frac": dom["gfpmul_s"] / peak_gfpmul,
                "peak_source": "measured in this run: rb_fq_mul_chain, %d threads x %d dependent-free Montgomery products" % (threads, 2 * iters),
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed
                # `ncu --set full` capture of this command (tools/gpu_profile_round.sh); null when it does not cover this batch size
                "traffic": ncu_traffic(dominant, B),
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, %s)" % NCU_TRAFFIC_FILE,
                "kernel_share_of_step": dom["ms"] / step_kernel_ms,
                "step_fp_mul": step_mul,
                "step_achieved": step_mul / ms_per_step / 1e6,
                "step_frac": step_mul / ms_per_step / 1e6 / peak_gfpmul,
                "per_kernel": per_kernel,
                "hbm": {"algorithmic_bytes_per_step": hbm_bytes, "achieved_gbs": hbm_bytes / ms_per_step / 1e6,
                        "peak_gbs": peaks.get("hbm_gbs"), "note": "reported to show HBM is not the limiter"},
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
            rs = line["cpu_baseline"]["restructured"]
            rs["hardware_factor_e2e"] = line["e2e"]["value"] / rs["value"]      # same algorithm, B200 vs the host's cores
        if world == 1 and not args.no_other_configs and not DISTINCT:
            line["other_configs"] = other_configs(local_rank, peak_gfpmul)
            line["other_configs"]["config_2_distinct"] = distinct_first
        print(json.dumps(line))
    rd.finalize()


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
    # second CPU mode: the RESTRUCTURED algorithm (oracle/ac17_fast.cpp: the algebra of the CUDA path -- folded policy
    # scalars, fixed-base window tables, one final exponentiation per decryption) on the same cores, so that the
    # GPU-vs-reference ratio splits into an algorithmic and a hardware factor by measurement
    import sys as _sys
    _sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rb_testutil as util
    fast = oracle.Ac17Fast(pk, m, pi)
    ct_idx, sk_idx = util.decrypt_lists(pruned, pi, names)
    fsample = 16 * sample
    frnd = [fr() + fr() for _ in range(fsample)]

    def one_fast(i):
        c0, c, cp = fast.encrypt(frnd[i], msg)
        assert oracle.Ac17Fast.decrypt(ct_idx, sk_idx, c0, c, cp, k0, k, kp) == msg
        return 1

    one_fast(0)
    t2 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        fdone = sum(ex.map(one_fast, range(fsample)))
    fdt = time.perf_counter() - t2
    t3 = time.perf_counter()
    for i in range(8):
        one_fast(i)
    fsingle = (time.perf_counter() - t3) / 8
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d round trips over %d threads (oracle/ac17.cpp reference-sequence restatement; not the Rust binary)" % (sample, cores),
            "single_thread_value": 1.0 / single,
            "restructured": {"value": fdone / fdt, "unit": UNIT, "cores": cores,
                             "sample": "%d round trips over %d threads, oracle/ac17_fast.cpp: the CUDA path's algebra on the CPU (byte-identical outputs)" % (fsample, cores),
                             "single_thread_value": 1.0 / fsingle,
                             "algorithm_factor": single / fsingle,
                             "algorithm_factor_note": "single-thread ratio (the multi-thread figure of the restructured mode is held back by the Python driver: ~10 ms calls, GIL-held buffer copies)"}}


if __name__ == "__main__":
    main()
