import sys, os, json
sys.path.insert(0, "/root/repo")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch, random
import oracle
from oracle import policy as opol
sys.path.insert(0, "/root/repo/tests")
import rb_testutil as util
from rb_testutil import u8
from rabe_b200.engine import Engine
rng = random.Random(1)
pk, msk = oracle.ac17_setup(util.rand_fr(rng, 9))
names = ["a%d" % i for i in range(8)]
tree = opol.parse(util.and_policy(names), opol.HUMAN)
m, pi, n2 = opol.calculate_msp(tree)
h_row, h_col = util.ac17_hashes(pi, n2)
k0, k, kp = oracle.ac17_cp_keygen(msk, names, util.rand_fr(rng, len(names) + 3))
ok, pruned = opol.calc_pruned(names, tree)
ct_idx, sk_idx = util.decrypt_lists(pruned, pi, names)
E = Engine(0); pkh = E.ac17_pk_load(u8(pk)); msp = E.msp_load(np.array(m, dtype=np.int8), u8(h_row), u8(h_col))
B = 2048
dev = torch.device("cuda", 0)
s = torch.from_numpy(u8(util.rand_fr(rng, 2 * B))).to(dev)
gt = E.gt_table(u8(pk[448:832]), 8)
msg = E.gt_pow_fixed(gt, torch.from_numpy(u8(util.rand_fr(rng, B))).to(dev))
c0, c, cp = E.ac17_cp_encrypt(pkh, msp, s, msg)
E.status()
D = [Engine(0) for _ in range(4)]
streams = [torch.cuda.Stream() for _ in D]
sk = []
for d, st in zip(D, streams):
    with torch.cuda.stream(st):
        d.use_torch_stream()
    sk.append(d.ac17_sk_load(u8(k0), u8(k), u8(kp)))
    d.set_g2_subgroup_check(False)
ci = torch.from_numpy(np.array(ct_idx, dtype=np.uint32).view(np.int32)).to(dev); si = torch.from_numpy(np.array(sk_idx, dtype=np.uint32).view(np.int32)).to(dev)
torch.cuda.synchronize()
for d in D: d.profile(True)
outs = [torch.empty(B * 384, dtype=torch.uint8, device=dev) for _ in D]
for kk in range(16):
    i = kk % 4
    with torch.cuda.stream(streams[i]):
        D[i].ac17_cp_decrypt_sk(sk[i], c0, c, cp, len(pi), ci, si, out=outs[i])
torch.cuda.synchronize()
for d in D:
    print({k_: v["launches"] for k_, v in d.profile_report().items()})
print(all(bool((o == msg).all().item()) for o in outs))
