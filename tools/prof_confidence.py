import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import pandora_b200
eng = pandora_b200.get_engine("cuda:0")
H, W, D = 4096, 4096, 256
g = torch.Generator(device="cuda").manual_seed(1)
cv = torch.randint(200, 460, (H, W, D), device="cuda", generator=g).float()
etas = np.arange(0.0, 0.7, 0.01); dr = np.arange(-(D - 1), 1).astype(np.float32)
for name, kw in (("ambiguity", {}), ("ambiguity+risk", {"risk": True})):
    for _ in range(2):
        eng.confidence(cv, etas, None, dr, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.confidence(cv, etas, None, dr, **kw); b.record(); torch.cuda.synchronize()
    print(name, round(a.elapsed_time(b), 2), "ms")
