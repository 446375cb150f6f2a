"""GPU parity tests at the BASELINE.json configurations themselves (round-1 verdict, "untested configs").

Every kernel instantiation the headline configurations run is compared with the CPU oracle here -- not with another
GPU kernel -- and every test asserts WHICH kernel family served the call (``pb200_last_path``), so a silent fall-back
to a slower path cannot pass:

* ``cbca_aggregate_reg_kernel<128 | 192 | 256>`` (C2 runs <192>) on short images with the full disparity range, and a
  2048-column x 192-disparity band cut out of the full-size C2 result;
* the fused Census -> SGM wavefront kernels at 4096 and 4144 columns x 256 disparities (C3's strip geometry: 147 / 148
  strips, 14 warps, partial last strip) directly against oracle.census -> oracle.sgm -> oracle.wta;
* C3's configuration through ``pandora_b200.run(cfg)`` (the plugin-level call) on a 4096-column band;
* at C3's full size, three independent implementations of the stage (one-column wavefront, two-column wavefront,
  four-launch packed schedule) must agree bit for bit.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pandora_b200

    return pandora_b200


def dev(eng, a):
    return eng.to_device(np.ascontiguousarray(a, dtype=np.float32))


def host(t):
    return t.detach().cpu().numpy()


def textured_pair(oracle, H, W, D):
    left, right, _ = oracle.synthetic_pair(H, W, D)
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


# ------------------------------------------------------------------------------------------------
# CBCA: the compile-time-D instantiations of the register kernel
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,D", [(40, 300, 192), (33, 200, 256), (36, 260, 128), (30, 90, 64)])
def test_cbca_register_kernel_fixed_d_vs_oracle(pb, oracle, H, W, D):
    eng = pb.get_engine("cuda:0")
    left, right = textured_pair(oracle, H, W, D)
    dmin = -(D - 1)
    cv, attrs = oracle.census_cost_volume(left, right, 5, dmin, 0)
    ref, _ = oracle.cbca_cost_volume(left, right, cv, 2, dmin, 5, 30.0, attrs["cmax"])
    got = eng.cbca(dev(eng, left), dev(eng, right), dev(eng, cv), 2, dmin, 5, 30.0)
    assert pb.last_path("cbca") == ("cbca_reg", D)
    np.testing.assert_array_equal(host(got), ref)          # integer costs: exact sums, one correctly rounded division


def test_c2_full_size_band_vs_oracle(pb, oracle):
    """C2 (2048 x 2048, Census 5x5 + CBCA, D = 192) at full size; a 40-row band of the result is compared with the
    oracle run on the band plus 12 rows of context on either side (CBCA reaches 4 rows through its arms and vertical
    sums, 1 more through the 3x3 median of the support image, 2 through the Census window: aggregation.cpp:28-221,
    cbca.py:217-295)."""
    H = W = 2048
    D = 192
    dmin = -(D - 1)
    left, right = textured_pair(oracle, H, W, D)
    pipe = pb.StereoPipeline(H, W, dmin, 0, "census", 5, cbca=(5, 30.0))
    disp = pipe.run_host(left, right).copy()
    assert pb.last_path("cbca") == ("cbca_reg", 192)
    r0, rows, ctx = 1000, 40, 12
    band = slice(r0 - ctx, r0 + rows + ctx)
    cv, attrs = oracle.census_cost_volume(left[band], right[band], 5, dmin, 0)
    ref, _ = oracle.cbca_cost_volume(left[band], right[band], cv, 2, dmin, 5, 30.0, attrs["cmax"])
    got = host(pipe.final_cv[r0:r0 + rows])
    np.testing.assert_array_equal(got, ref[ctx:ctx + rows])
    exp, _ = oracle.wta(ref, np.arange(dmin, 1))
    np.testing.assert_array_equal(disp[r0:r0 + rows], exp[ctx:ctx + rows])


# ------------------------------------------------------------------------------------------------
# fused Census -> SGM at C3's strip geometry, directly against the oracle chain
# ------------------------------------------------------------------------------------------------
def oracle_chain(oracle, left, right, dmin, dmax, p1=8, p2=32):
    cv, attrs = oracle.census_cost_volume(left, right, 5, dmin, dmax)
    S = oracle.sgm_cost_volume(cv, p1, p2, cmax=attrs["cmax"])
    disp, inv = oracle.wta(S, np.arange(dmin, dmax + 1))
    return S, disp, inv


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("H,W", [(48, 4096), (8, 4144), (12, 4095)])
def test_fused_census_sgm_c3_geometry_vs_oracle(pb, oracle, H, W, kernel):
    eng = pb.get_engine("cuda:0")
    D, dmin = 256, -255
    left, right = textured_pair(oracle, H, W, D)
    S, disp, inv = oracle_chain(oracle, left, right, dmin, 0)
    with pb.option("sgm.wave_kernel", kernel):
        out = eng.census_sgm(dev(eng, left), dev(eng, right), 5, dmin, 0, 8, 32)
    assert out is not None, "C3's geometry must take the fused stage"
    assert pb.last_path("sgm")[0] == ("sgm_wave1_census" if kernel == 1 else "sgm_wave2_census")
    np.testing.assert_array_equal(host(out[0]), S)
    np.testing.assert_array_equal(host(out[1]), disp)
    np.testing.assert_array_equal(host(out[2]).astype(bool), inv)


def test_c3_through_the_plugin_call_vs_oracle(pb, oracle):
    """C3's configuration through ``pandora_b200.run(cfg)`` -- the step classes behind the reference's plugin API -- on a
    32-row band of the 4096-column pair: disparity map, validity mask and cost volume against the oracle chain."""
    H, W, D, dmin = 32, 4096, 256, -255
    left, right = textured_pair(oracle, H, W, D)
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5, "subpix": 1},
                        "optimization": {"optimization_method": "sgm", "penalty": {"P1": 8, "P2": 32}},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": -9999}}}
    dl = pb.create_image_dataset(left, disparity=[dmin, 0])
    dr = pb.create_image_dataset(right)
    disp, cv = pb.run(dl, dr, cfg)
    assert pb.last_path("sgm")[0] in ("sgm_wave1_census", "sgm_wave2_census")
    S, exp, inv = oracle_chain(oracle, left, right, dmin, 0)
    np.testing.assert_array_equal(np.asarray(disp["disparity_map"].data), exp)
    np.testing.assert_array_equal(np.asarray(cv["cost_volume"].data), S)
    mask = oracle.validity_mask(H, W, dmin, 0, 2)
    oracle.cv_masked(oracle.census_cost_volume(left, right, 5, dmin, 0)[0], mask, 2)
    np.testing.assert_array_equal(np.asarray(disp["validity_mask"].data), oracle.wta_validity_mask(mask, inv))


def test_c3_full_size_three_implementations_agree(pb, oracle):
    """4096 x 4096 x 256 is out of the oracle's reach (hours); the one-column wavefront kernels, the two-column ones and
    the four-launch packed schedule on the materialised Census volume share no recurrence code path beyond `nstep`, and
    must produce the same 17 GB volume and the same disparity map."""
    import torch

    eng = pb.get_engine("cuda:0")
    H = W = 4096
    D, dmin = 256, -255
    left, right = textured_pair(oracle, H, W, D)
    dl, dr = dev(eng, left), dev(eng, right)
    outs = []
    for kernel in (1, 2):
        with pb.option("sgm.wave_kernel", kernel):
            out = eng.census_sgm(dl, dr, 5, dmin, 0, 8, 32)
        assert out is not None
        assert pb.last_path("sgm")[0] == ("sgm_wave1_census" if kernel == 1 else "sgm_wave2_census")
        outs.append((out[1].clone(), out[2].clone(), torch.nan_to_num(out[0], nan=-7.0).to(torch.float16)))   # sums < 2048: exact in fp16
        del out
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    with pb.option("sgm.no_wave", 1):
        cvol = eng.census(dl, dr, 5, dmin, 0)
        S, disp, flags = eng.sgm(cvol, 8, 32, 58.0, fuse_wta=True, dmin=dmin)
    assert pb.last_path("sgm")[0] == "sgm_packed4"
    assert torch.equal(disp, outs[0][0]) and torch.equal(flags, outs[0][1])
    assert torch.equal(torch.nan_to_num(S, nan=-7.0).to(torch.float16), outs[0][2])
