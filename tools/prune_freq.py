import sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
from gfnerf_b200.engine import GFNeRFEngine
from tests.helpers import make_sampler
rig = bench.load_rig(); dev = torch.device("cuda", 0)
s = make_sampler(rig, mode=0, device=dev); s.ray_march_fineness_decay_end_iter_ = 0.0; s.ray_march_fineness_ = 1.0
s.generator = torch.Generator(device=dev).manual_seed(1234)
eng = GFNeRFEngine(s, log2_table_size=19, num_images=rig["c2w"].shape[0], seed=0)
host = bench.make_batches(rig, 8192, 8, seed=1234)
res = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in host]
def valid():
    n = s.tree_nodes_gpu_.view(-1, 128)[:, 96:104].contiguous().view(torch.int64).view(-1)
    return int((n >= 0).sum().item())
prev = valid(); changes = []
for i in range(300):
    o, d, cam, tgt = res[i % 8]
    eng.train_step(o, d, tgt, cam, next_rays=res[(i + 1) % 8][:2])
    v = valid()
    changes.append(prev - v); prev = v
c = np.array(changes)
print("valid leaves at end", prev, "steps with a pruning event:", int((c != 0).sum()), "of", len(c))
print("events in steps 0-50:", int((c[:50] != 0).sum()), " 50-150:", int((c[50:150] != 0).sum()), " 150-300:", int((c[150:] != 0).sum()))
print("pruned per event:", c[c != 0][:40])
