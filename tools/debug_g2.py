import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from rabe_b200.engine import Engine
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
fr = lambda x: int(x % R).to_bytes(32, "big")
u8 = lambda b: np.frombuffer(bytes(b), dtype=np.uint8).copy()
eng = Engine(0)
h = oracle.g2_generator()
rng = random.Random(3)
ks = [2, 3, 4, 5, 7, 8, 15, 16, 255, 256, 65535, 1 << 20, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 100, (1 << 128) + 12345, 1 << 200, 1 << 253, R - 1] + [rng.randrange(1 << b) for b in (10, 20, 33, 64, 65, 100, 130, 200, 250, 253)]
out = eng.g2_mul_var(u8(h * len(ks)), u8(b"".join(fr(k) for k in ks))).tobytes()
for i, k in enumerate(ks):
    ok = out[128 * i:128 * i + 128] == oracle.g2_mul(h, fr(k))
    print(hex(k), "OK" if ok else "FAIL")
g = oracle.g1_generator()
out = eng.g1_mul_var(u8(g * len(ks)), u8(b"".join(fr(k) for k in ks))).tobytes()
print("g1 fails:", [hex(k) for i, k in enumerate(ks) if out[64 * i:64 * i + 64] != oracle.g1_mul(g, fr(k))])
