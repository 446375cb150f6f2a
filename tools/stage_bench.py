#!/usr/bin/env python3
"""Per-kernel timing (CUDA events, min of reps) at the BASELINE.json configurations, with the algorithmic-byte
roofline fraction of each stage.  usage: python tools/stage_bench.py [--json out.json]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402

PEAK = 6482.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
eng = pandora_b200.get_engine("cuda:0")
out = []


def timeit(name, fn, alg_bytes, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts)
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    rec = {"stage": name, "ms": round(ms, 4), "algorithmic_GB": round(alg_bytes / 1e9, 3), "achieved_GBs": round(gbs, 1), "frac_of_measured": round(gbs / PEAK, 3)}
    out.append(rec)
    print(json.dumps(rec), flush=True)


def pair(H, W, D):
    l, r, _ = synthetic_pair(H, W, D)
    return eng.to_device(l), eng.to_device(r)


# C0-like: cones size, SAD w5, D=64
H, W, D = 375, 450, 64
l, r = pair(H, W, D)
cv = eng.empty((H, W, D))
timeit("C0 sad w5 375x450x64", lambda: eng.sad_ssd(l, r, 5, -(D - 1), 0, out=cv), (4 * D + 8) * H * W)
timeit("C0 wta 375x450x64", lambda: eng.wta(cv, -(D - 1)), (4 * D + 6) * H * W)
timeit("C0 zncc w5 375x450x64", lambda: eng.zncc(l, r, 5, -(D - 1), 0, out=cv), (4 * D + 8) * H * W)

# C1: 1024x1024 census D=128 + WTA (fused)
H, W, D = 1024, 1024, 128
l, r = pair(H, W, D)
cv = eng.empty((H, W, D))
timeit("C1 census fill 1024x1024x128", lambda: eng.census(l, r, 5, -(D - 1), 0, out=cv), (4 * D + 8) * H * W)
p1 = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5)
timeit("C1 census+fused WTA pipeline", lambda: p1.run_device(l, r), (4 * D + 12) * H * W)
timeit("C1 wta standalone", lambda: eng.wta(cv, -(D - 1)), (4 * D + 6) * H * W)
timeit("C1 sad w5 1024x1024x128", lambda: eng.sad_ssd(l, r, 5, -(D - 1), 0, out=cv), (4 * D + 8) * H * W)
timeit("C1 zncc w5 1024x1024x128", lambda: eng.zncc(l, r, 5, -(D - 1), 0, out=cv), (4 * D + 8) * H * W)

# C2: 2048x2048 census + CBCA D=192
H, W, D = 2048, 2048, 192
l, r = pair(H, W, D)
cv = eng.empty((H, W, D))
cv2 = eng.empty((H, W, D))
timeit("C2 census fill 2048x2048x192", lambda: eng.census(l, r, 5, -(D - 1), 0, out=cv), (4 * D + 8) * H * W)
timeit("C2 cbca supports (2x median3 + 2x cross_support)", lambda: eng.cbca_supports(l, r, 2, 5, 30.0), 2 * (4 + 4 + 4 + 8) * H * W)
timeit("C2 cbca (supports + aggregate) 2048x2048x192", lambda: eng.cbca(l, r, cv, 2, -(D - 1), 5, 30.0, out=cv2), (8 * D + 24) * H * W)
cl2, cr2 = eng.cbca_supports(l, r, 2, 5, 30.0)
timeit("C2 cbca aggregate only (register kernel)", lambda: pandora_b200._native.check(eng.lib.pb200_cbca_aggregate(
    cv.data_ptr(), cv2.data_ptr(), H, W, D, -(D - 1), 2, cl2.data_ptr(), cr2.data_ptr(), 5, torch.cuda.current_stream().cuda_stream)),
    (8 * D + 24) * H * W)
timeit("C2 wta 2048x2048x192", lambda: eng.wta(cv2, -(D - 1)), (4 * D + 6) * H * W)
p2 = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, cbca=(5, 30.0))
timeit("C2 pipeline census+cbca+wta", lambda: p2.run_device(l, r), (12 * D + 36) * H * W)
del cv, cv2, p1, p2
torch.cuda.empty_cache()

# C3: 4096x4096 census + SGM D=256
H, W, D = 4096, 4096, 256
l, r = pair(H, W, D)
p3 = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, sgm=(8.0, 32.0))
timeit("C3 census fill 4096x4096x256", lambda: eng.census(l, r, 5, -(D - 1), 0, out=p3.cv_a), (4 * D + 8) * H * W)
timeit("C3 sgm 8-path + fused WTA", lambda: eng.sgm(p3.cv_a, 8.0, 32.0, 58.0, out=p3.cv_b, fuse_wta=True, dmin=-(D - 1), disp=p3.disp, flags=p3.flags), 8 * D * H * W)
timeit("C3 fused census+sgm+wta stage (pb200_census_sgm incl. the two transforms)",
       lambda: eng.census_sgm(l, r, 5, -(D - 1), 0, 8.0, 32.0, out=p3.cv_b, disp=p3.disp, flags=p3.flags), 8 * D * H * W)
timeit("C3 pipeline census+sgm+wta (StereoPipeline, fused stage)", lambda: p3.run_device(l, r), (12 * D + 12) * H * W)
timeit("C3 wta standalone 4096x4096x256", lambda: eng.wta(p3.cv_b, -(D - 1)), (4 * D + 6) * H * W)
timeit("C3 reverse_cost_volume", lambda: eng.reverse_cost_volume(p3.cv_a, 0), 8 * D * H * W, reps=2)
# next rows on the C3 SGM result: algorithmic bytes = one read of the volume (+ O(H*W) maps)
dmin3 = -(D - 1)
timeit("C3 wta_right (right map from the left volume)", lambda: eng.wta_right(p3.cv_b, 0), (4 * D + 5) * H * W)
timeit("C3 reverse_cost_volume + wta (what wta_right replaces)", lambda: eng.wta(eng.reverse_cost_volume(p3.cv_b, 0), 0), (4 * D + 5) * H * W, reps=2)
timeit("C3 cv_masked pass (all-NaN detection, no masks)", lambda: eng.cv_masked(p3.cv_b, dmin3), (4 * D + 1) * H * W)
msk = (torch.rand((H, W), device="cuda") < 0.02).to(torch.int16).cpu().numpy()
fl3 = eng.mask_flags(msk, 0, 1, 5)
d_msk = torch.from_numpy(msk).to("cuda")
timeit("C3 mask_flags (5x5 no_data dilation, mask resident)", lambda: eng.mask_flags(d_msk, 0, 1, 5), 3 * H * W, reps=3)
timeit("C3 mask_flags incl. the pageable H2D of the 33 MB mask", lambda: eng.mask_flags(msk, 0, 1, 5), 3 * H * W, reps=2)
vm3 = eng.validity_mask_init(H, W, dmin3, 0, 2)
timeit("C3 validity_mask_masks (left + right msk)", lambda: eng.validity_mask_masks(vm3, dmin3, 0, 2, fl3, fl3), 6 * H * W)
disp3, flags3 = eng.wta(p3.cv_b, dmin3)
mask3 = eng.validity_mask(H, W, dmin3, 0, 2, flags3, wta_invalidate=True, mask=eng.validity_mask(H, W, dmin3, 0, 2, flags3))
rdisp3, _ = eng.wta_right(p3.cv_b, 0)
timeit("C3 refinement vfit", lambda: eng.refinement(p3.cv_b, disp3.clone(), mask3.clone(), dmin3, 0, 1, False, "vfit"), (3 * 32 + 4 + 2 + 4 + 2 + 4) * H * W)
timeit("C3 cross_checking", lambda: eng.cross_checking(disp3, mask3.clone(), rdisp3, 1.0, dmin3, 0, 2), (4 + 4 + 2 + 2 + 4) * H * W)
etas3 = np.arange(0.0, 0.7, 0.01)
dr3 = np.arange(dmin3, 1).astype(np.float32)
timeit("C3 ambiguity (70 etas; 2 reads of the volume)", lambda: eng.confidence(p3.cv_b, etas3, None, dr3), (8 * D + 8) * H * W, reps=2)
timeit("C3 ambiguity + risk (70 etas)", lambda: eng.confidence(p3.cv_b, etas3, None, dr3, risk=True), (8 * D + 24) * H * W, reps=2)
# float costs: the float SGM kernels (no packed path)
p3.cv_a.mul_(0.37)
timeit("C3 sgm float costs (float kernels)", lambda: eng.sgm(p3.cv_a, 0.3, 1.7, 12.0, out=p3.cv_b), 8 * D * H * W, reps=2)
if len(sys.argv) > 2 and sys.argv[1] == "--json":
    json.dump(out, open(sys.argv[2], "w"), indent=1)
