import os, sys, random, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rabe_b200.engine import Engine
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
eng = Engine(0); s = torch.cuda.Stream(); torch.cuda.set_stream(s); eng.use_torch_stream()
rng = random.Random(1)
base = np.frombuffer(b"".join(rng.randrange(P).to_bytes(32, "big") for _ in range(1024)), dtype=np.uint8)
for threads in (148 * 128, 148 * 256, 148 * 512, 148 * 2048):
    a = torch.from_numpy(np.tile(base, threads // 1024 + 1)[:32 * threads].copy()).cuda()
    for ilp, code, nmul in ((1, -2000, 1), (2, 2000, 2), (4, (1 << 20) + 2000, 4)):
        eng.fq_mul_chain(a, a, code); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); eng.fq_mul_chain(a, a, code); e1.record(s); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"threads": threads, "warps_per_smsp": threads / 32 / 592, "ilp": ilp, "ms": round(best, 3), "gmul_s": round(threads * 2000 * nmul / best / 1e6, 1)}))
