"""GPU parity tests of the fused Census -> SGM -> WTA stage (``pb200_census_sgm``: the first wavefront pass computes
its Hamming costs from the census descriptors, the float32 Census volume is never written).

Bar: bit-exact against the oracle chain census_cost_volume -> sgm_cost_volume -> wta, and bit-identical to the two
separate C-ABI calls (``pb200_census_cost_volume`` + ``pb200_sgm``) at sizes the oracle would take minutes for.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pandora_b200

    return pandora_b200.get_engine("cuda:0")


def dev(eng, a):
    return eng.to_device(np.ascontiguousarray(a, dtype=np.float32))


def host(t):
    return t.detach().cpu().numpy()


def pair(seed, H, W, levels=256):
    g = np.random.default_rng(seed)
    base = g.integers(0, levels, (H, W + 8)).astype(np.float32)
    left = base[:, 4:4 + W].copy()
    right = np.roll(base, 3, axis=1)[:, 4:4 + W].copy()
    right[g.random((H, W)) < 0.2] += 1.0                       # not a pure shift: every disparity has a non-trivial cost
    return left, right


def oracle_chain(oracle, left, right, w, dmin, dmax, p1, p2, over):
    cv, attrs = oracle.census_cost_volume(left, right, w, dmin, dmax)
    assert attrs["cmax"] == w * w
    S = oracle.sgm_cost_volume(cv, p1, p2, cmax=attrs["cmax"], overcounting=over)
    disp, inv = oracle.wta(S, np.arange(dmin, dmax + 1))
    return S, disp, inv


@pytest.mark.parametrize("H,W,D,dmin", [
    (21, 40, 64, -63), (21, 40, 64, -20), (16, 90, 64, 5), (9, 3, 64, -63), (1, 70, 64, -10), (30, 1, 64, -63),
    (33, 333, 128, -127), (18, 300, 128, -64), (12, 150, 128, 0),
    (24, 600, 256, -255), (10, 700, 256, -100), (7, 260, 256, 3),
    (26, 500, 192, -191), (12, 4100, 192, -100), (9, 130, 192, 4),       # D = 192 (C2's disparity count): three registers per lane
])
@pytest.mark.parametrize("w", [5, 3])
def test_census_sgm_fused_vs_oracle(eng, oracle, H, W, D, dmin, w):
    """All three disparity counts of the packed path, both one-word windows, ranges that leave the image on either side
    (NaN cells from the descriptor flags and from columns outside the descriptor row), images narrower than a strip."""
    left, right = pair(H * 1000 + W + D + w, H, W, levels=9)
    dmax = dmin + D - 1
    S, disp, inv = oracle_chain(oracle, left, right, w, dmin, dmax, 8, 32, False)
    out = eng.census_sgm(dev(eng, left), dev(eng, right), w, dmin, dmax, 8, 32)
    assert out is not None, "eligible configuration was not fused"
    got, gdisp, gflags = out
    np.testing.assert_array_equal(host(got), S)
    np.testing.assert_array_equal(host(gdisp), disp)
    np.testing.assert_array_equal(host(gflags).astype(bool), inv)


@pytest.mark.parametrize("over,p1,p2", [(True, 8, 32), (False, 10, 120), (True, 3, 200), (False, 1, 1)])
@pytest.mark.parametrize("H,W,D", [(14, 200, 64), (11, 310, 128), (9, 420, 256), (10, 350, 192)])
def test_census_sgm_fused_penalties_and_overcounting(eng, oracle, H, W, D, over, p1, p2):
    """Byte storage tier (cost + P2 <= 127) and 16-bit tier, overcounting, no fused WTA."""
    left, right = pair(p2 * 7 + D, H, W)
    dmin, dmax = -(D - 1) + 17, 17
    S, _, _ = oracle_chain(oracle, left, right, 5, dmin, dmax, p1, p2, over)
    out = eng.census_sgm(dev(eng, left), dev(eng, right), 5, dmin, dmax, p1, p2, overcounting=over, fuse_wta=False)
    assert out is not None
    np.testing.assert_array_equal(host(out[0]), S)


@pytest.mark.parametrize("H,W,D", [(6, 4096, 256), (5, 4095, 128), (9, 3001, 64), (40, 1500, 256), (3, 4144, 256)])
def test_census_sgm_fused_wide_images_equal_separate_calls(eng, H, W, D):
    """Strips of 14 warps, partial last strips, the widest image one wave takes: identical bits to census + sgm."""
    import torch

    left, right = pair(W + D, H, W)
    dl, dr = dev(eng, left), dev(eng, right)
    dmin, dmax = -(D - 1), 0
    S_ref, disp_ref, flags_ref = eng.sgm(eng.census(dl, dr, 5, dmin, dmax), 8, 32, 58.0, fuse_wta=True, dmin=dmin)
    out = eng.census_sgm(dl, dr, 5, dmin, dmax, 8, 32)
    assert out is not None
    assert torch.equal(torch.nan_to_num(out[0], nan=-7.0), torch.nan_to_num(S_ref, nan=-7.0))
    assert torch.equal(out[1], disp_ref)
    assert torch.equal(out[2], flags_ref)


def test_census_sgm_fused_tall_image_repeatable(eng):
    """Many rows (staging ring wrap-around, one-row-ahead pixel A at the last row) and two runs with the same bits."""
    import torch

    H, W, D = 700, 900, 256
    left, right = pair(77, H, W)
    dl, dr = dev(eng, left), dev(eng, right)
    S_ref, disp_ref, _ = eng.sgm(eng.census(dl, dr, 5, -255, 0), 8, 32, 58.0, fuse_wta=True, dmin=-255)
    for _ in range(2):
        out = eng.census_sgm(dl, dr, 5, -255, 0, 8, 32)
        assert out is not None
        assert torch.equal(torch.nan_to_num(out[0], nan=-7.0), torch.nan_to_num(S_ref, nan=-7.0))
        assert torch.equal(out[1], disp_ref)


@pytest.mark.parametrize("w,D", [(7, 64), (5, 100), (5, 300), (13, 256)])
def test_census_sgm_not_eligible_returns_none(eng, w, D):
    """Two-word descriptors and disparity counts outside {64, 128, 192, 256}: nothing is computed, the caller falls back."""
    left, right = pair(1, 30, 80)
    assert eng.census_sgm(dev(eng, left), dev(eng, right), w, -(D - 1), 0, 8, 32) is None


@pytest.mark.parametrize("shape", [(300, 257, 64), (520, 600, 256)])
def test_stereo_pipeline_fused_equals_unfused(eng, shape):
    """StereoPipeline with and without the fused stage: same disparity map and SGM volume, device and host entry."""
    import torch

    import pandora_b200

    H, W, D = shape
    left, right = pair(D, H, W)
    ref_pipe = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, sgm=(8, 32), fuse_census_sgm=False)
    pipe = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, sgm=(8, 32), fuse_census_sgm=True)
    dl, dr = eng.to_device(left), eng.to_device(right)
    ref = ref_pipe.run_device(dl, dr).clone()
    assert not ref_pipe.fused_ran
    got = pipe.run_device(dl, dr)
    assert pipe.fused_ran
    assert torch.equal(got, ref)
    assert torch.equal(torch.nan_to_num(pipe.final_cv, nan=-7.0), torch.nan_to_num(ref_pipe.final_cv, nan=-7.0))
    assert torch.equal(pipe.validity_mask(), ref_pipe.validity_mask())
    pipe.disp.fill_(0.0)
    np.testing.assert_array_equal(pipe.run_host(left, right), host(ref))
    # a configuration the fused stage does not take goes through the separate steps
    other = pandora_b200.StereoPipeline(40, 64, -19, 0, "census", 5, sgm=(8, 32), fuse_census_sgm=True)
    l2, r2 = pair(3, 40, 64)
    other.run_device(eng.to_device(l2), eng.to_device(r2))
    assert not other.fused_ran


def test_disparity_host_fused(eng, oracle):
    """pb200_disparity_host with the fused stage switched on: same outputs as the oracle chain."""
    from pandora_b200 import _native

    lib = _native.load()
    left, right = pair(5, 36, 150)
    H, W = left.shape
    dmin, dmax = -63, 0
    S, exp, inv = oracle_chain(oracle, left, right, 5, dmin, dmax, 8, 32, False)
    mask = oracle.validity_mask(H, W, dmin, dmax, 2)
    oracle.cv_masked(oracle.census_cost_volume(left, right, 5, dmin, dmax)[0], mask, 2)     # the all-NaN bits of the cost step
    disp = np.empty((H, W), dtype=np.float32)
    vm = np.empty((H, W), dtype=np.uint16)
    out = np.empty((H, W, 64), dtype=np.float32)
    with _native.option("fuse_census_sgm", 1):
        before = _native.kernel_launches()
        _native.check(lib.pb200_disparity_host(left.ctypes.data, right.ctypes.data, H, W, 0, 5, dmin, dmax, 0, 30.0, 8.0, 32.0, 0, -9999.0,
                                               disp.ctypes.data, vm.ctypes.data, out.ctypes.data))
        launched = _native.kernel_launches() - before
    np.testing.assert_array_equal(out, S)
    np.testing.assert_array_equal(disp, exp)
    np.testing.assert_array_equal(vm, oracle.wta_validity_mask(mask, inv))
    assert launched <= 7, launched        # 2 transforms + 2 wavefront passes + 3 validity-mask kernels: no Census fill


@pytest.mark.parametrize("cfg", [dict(sgm=(8, 32)), dict(sgm=(8, 32), fuse_census_sgm=False), dict(), dict(cbca=(5, 30.0))])
def test_stream_of_pairs_equals_single_calls(eng, cfg):
    """submit_host / result_host: a stream of different pairs, two in flight, pinned and pageable inputs -- every
    disparity map equals the synchronous device run of the same pair (no buffer of a pair in flight is reused early)."""
    import torch

    import pandora_b200

    H, W, D = 300, 400, 64
    pipe = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, **cfg)
    pairs = [pair(100 + i, H, W) for i in range(5)]
    refs = [pipe.run_device(eng.to_device(l), eng.to_device(r)).cpu().numpy().copy() for l, r in pairs]
    inputs = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) if i % 2 == 0 else (l, r) for i, (l, r) in enumerate(pairs)]
    prev, got = None, {}
    for i, (l, r) in enumerate(inputs):
        tk = pipe.submit_host(l, r)
        assert tk == i
        if prev is not None:
            got[prev] = pipe.result_host(prev).copy()
        prev = tk
    got[prev] = pipe.result_host(prev).copy()
    for i in range(5):
        np.testing.assert_array_equal(got[i], refs[i])
    with pytest.raises(ValueError):
        pipe.result_host(1)                                   # no longer in flight
    # the synchronous entry still works on the same pipeline object
    np.testing.assert_array_equal(pipe.run_host(*pairs[2]), refs[2])


def test_plugin_steps_fuse_census_into_sgm(eng, oracle):
    """Through the step classes (pandora_b200.run): the Census step leaves a deferred volume, cv_masked gets its all-NaN
    pixels from the geometry, and a directly following SGM step runs the fused kernels -- same datasets as with the
    separate steps and as the oracle chain, without a Census fill launch; anything that reads the volume in between
    (here: CBCA) computes it."""
    import pandora_b200 as pb

    left, right = pair(21, 90, 200)
    dl = pb.create_image_dataset(left, disparity=[-63, 0])
    dr = pb.create_image_dataset(right)
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5},
                        "optimization": {"optimization_method": "sgm", "penalty": {"P1": 8, "P2": 32}},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": -9999}}}
    before = pb.kernel_launches()
    disp, cv = pb.run(dl, dr, cfg)
    fused_launches = pb.kernel_launches() - before
    S, exp, inv = oracle_chain(oracle, left, right, 5, -63, 0, 8, 32, False)
    np.testing.assert_array_equal(disp["disparity_map"].data, exp)
    np.testing.assert_array_equal(cv["cost_volume"].data, S)
    with pb.option("fuse_census_sgm", 0):
        before = pb.kernel_launches()
        disp2, cv2 = pb.run(dl, dr, cfg)
        plain_launches = pb.kernel_launches() - before
    np.testing.assert_array_equal(disp2["disparity_map"].data, disp["disparity_map"].data)
    np.testing.assert_array_equal(np.asarray(disp2["validity_mask"].data), np.asarray(disp["validity_mask"].data))
    np.testing.assert_array_equal(cv2["cost_volume"].data, cv["cost_volume"].data)
    assert fused_launches < plain_launches, (fused_launches, plain_launches)
    # a reader between the two steps (CBCA) gets the real Census volume
    cfg2 = {"pipeline": {"matching_cost": cfg["pipeline"]["matching_cost"], "aggregation": {"aggregation_method": "cbca"},
                         "optimization": cfg["pipeline"]["optimization"], "disparity": cfg["pipeline"]["disparity"]}}
    disp3, _ = pb.run(dl, dr, cfg2)
    ref, attrs = oracle.census_cost_volume(left, right, 5, -63, 0)
    ref, cmax = oracle.cbca_cost_volume(left, right, ref, 2, -63, 5, 30.0, attrs["cmax"])
    ref = oracle.sgm_cost_volume(ref, 8, 32, cmax=cmax)
    np.testing.assert_array_equal(disp3["disparity_map"].data, oracle.wta(ref, np.arange(-63, 1))[0])


# ---- batches of pairs through one wave (pb200_census_sgm_batch) ---------------------------------------------------------------
@pytest.mark.parametrize("n,H,W,D,dmin,w,p2", [(3, 40, 200, 64, -63, 5, 32), (2, 33, 4100, 256, -255, 5, 32), (4, 24, 700, 128, -100, 3, 32),
                                               (2, 50, 333, 256, -200, 5, 120), (5, 4, 96, 64, -20, 5, 32), (3, 20, 300, 192, -150, 5, 32)])
def test_census_sgm_batch_equals_single_calls(eng, oracle, n, H, W, D, dmin, w, p2):
    """A batch stacked into one wave: every image's volume, disparity map and all-NaN flags equal its own pb200_census_sgm
    call bit for bit (different images in the batch; paths restart at every image's first row in both passes); the first
    image is also checked against the oracle chain."""
    import torch

    pairs = [oracle.synthetic_pair(H, W, D, seed=100 + 7 * i)[:2] for i in range(n)]
    left = torch.stack([eng.to_device(p[0]) for p in pairs]).contiguous()
    right = torch.stack([eng.to_device(p[1]) for p in pairs]).contiguous()
    got = eng.census_sgm_batch(left, right, w, dmin, dmin + D - 1, 8, p2)
    assert got is not None
    for i in range(n):
        one = eng.census_sgm(left[i].contiguous(), right[i].contiguous(), w, dmin, dmin + D - 1, 8, p2)
        assert one is not None
        assert torch.equal(torch.nan_to_num(got[0][i], nan=-7.0), torch.nan_to_num(one[0], nan=-7.0)), f"volume of image {i}"
        assert torch.equal(got[1][i], one[1]) and torch.equal(got[2][i], one[2]), f"maps of image {i}"
    cv, attrs = oracle.census_cost_volume(pairs[0][0], pairs[0][1], w, dmin, dmin + D - 1)
    ref = oracle.sgm_cost_volume(cv, 8, p2, cmax=attrs["cmax"])
    np.testing.assert_array_equal(got[0][0].cpu().numpy(), ref)


def test_stereo_pipeline_batch(eng, oracle):
    import torch

    import pandora_b200

    n, H, W, D = 3, 48, 256, 128
    pairs = [oracle.synthetic_pair(H, W, D, seed=5 + i)[:2] for i in range(n)]
    pipe = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, sgm=(8, 32), device="cuda:0")
    left = torch.stack([eng.to_device(p[0]) for p in pairs]).contiguous()
    right = torch.stack([eng.to_device(p[1]) for p in pairs]).contiguous()
    disp = pipe.run_device_batch(left, right).clone()
    assert pipe.batched_ran
    for i in range(n):
        assert torch.equal(disp[i], pipe.run_device(left[i].contiguous(), right[i].contiguous()))
    with pandora_b200.option("sgm.wave_kernel", 1):            # the one-column kernels have no batched entry here: per-pair fall-back
        disp2 = pipe.run_device_batch(left, right)
        assert not pipe.batched_ran and torch.equal(disp2, disp)


def test_stream_of_batches_equals_single_calls(eng, oracle):
    """submit_host_batch / result_host_batch (batches of different sizes interleaved with single submissions, two in flight):
    every pair's map equals its own blocking call."""
    import pandora_b200

    H, W, D = 40, 200, 64
    pipe = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, sgm=(8, 32), device="cuda:0")
    pairs = [oracle.synthetic_pair(H, W, D, seed=300 + i)[:2] for i in range(9)]
    want = [pipe.run_host(l, r).copy() for l, r in pairs]
    plan = [(0, 4), (4, 1), (5, 3), (8, 1)]                     # (first pair, size)
    got, prev = {}, None

    def collect(p):
        a, n, tk, batch = p
        res = pipe.result_host_batch(tk) if batch else pipe.result_host(tk)[None]
        for i in range(n):
            got[a + i] = res[i].copy()

    for a, n in plan:
        ls, rs = [p[0] for p in pairs[a:a + n]], [p[1] for p in pairs[a:a + n]]
        tk = (a, n, pipe.submit_host_batch(ls, rs), True) if n > 1 else (a, n, pipe.submit_host(ls[0], rs[0]), False)
        if prev is not None:
            collect(prev)
        prev = tk
    collect(prev)
    for i in range(9):
        np.testing.assert_array_equal(got[i], want[i])
