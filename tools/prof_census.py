"""Census fill at C3, optionally with a pinned tile size: python tools/prof_census.py [tile]  (library option census.tile)"""
import sys, os, torch
sys.path.insert(0, os.getcwd())
import pandora_b200
from pandora_b200.synthetic import synthetic_pair
eng = pandora_b200.get_engine("cuda:0")
tile = int(sys.argv[1]) if len(sys.argv) > 1 else -1
pandora_b200.set_option("census.tile", tile)
H, W, D = 4096, 4096, 256
l, r, _ = synthetic_pair(H, W, D)
l, r = eng.to_device(l), eng.to_device(r)
cv = eng.empty((H, W, D))
for _ in range(3): eng.census(l, r, 5, -(D - 1), 0, out=cv)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.census(l, r, 5, -(D - 1), 0, out=cv); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("census.tile =", tile, "census C3", round(min(ts), 3), "ms")
