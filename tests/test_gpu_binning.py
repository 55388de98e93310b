"""-m gpu: projection / AABB / tile binning on the GPU against the CPU oracle.

Integer outputs (tile counts given centres, cumsum, keys, sorted keys, sorted ids, tile ranges) must be
BIT-EXACT.  Float outputs (centres, extents, depths) are compared with rtol 1e-5 / atol 1e-4 pixel.
"""
import numpy as np
import pytest
import torch

import oracle
from gstex_cuda_b200 import get_aabb_2d as A
from gstex_cuda_b200 import utils as U
from gstex_cuda_b200 import cuda as _C
from gstex_cuda_b200.scenes import random_small_scene, synthetic_scene
from gpu_util import DEV, to_np, bin_cuda, assert_close_frac

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,W,H,bw", [(100, 64, 64, 16), (1000, 200, 120, 16), (777, 130, 70, 8), (5000, 256, 256, 16)])
def test_aabb_depth_and_counts(n, W, H, bw):
    s = random_small_scene(n, W, H, seed=n, device=DEV)
    # push a few Gaussians behind / onto the near plane (clipped path, get_aabb_2d.cu:81-84)
    s["means"][: max(1, n // 50), 2] = -8.0 + torch.linspace(-0.5, 0.02, max(1, n // 50), device=DEV)
    intr = s["intrins"]
    pix, depths = A.project_points(s["means"], s["viewmat"], intr)
    centers, extents = A.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], intr)
    pix_o, depths_o = oracle.project_points(to_np(s["means"]), to_np(s["viewmat"]), intr)
    c_o, e_o = oracle.get_aabb_2d(to_np(s["means"]), to_np(s["scales"]), 1.0, to_np(s["quats"]), to_np(s["viewmat"]), intr)
    assert_close_frac("depths", to_np(depths), depths_o, 1e-6, 1e-6)
    assert_close_frac("pix", to_np(pix), pix_o, 1e-5, 1e-3)
    assert_close_frac("centers", to_np(centers), c_o, 1e-5, 1e-3)
    assert_close_frac("extents", to_np(extents), e_o, 1e-5, 1e-3)
    assert np.array_equal(to_np(extents) == 0, e_o == 0)  # same Gaussians clipped
    # tile counts from the SAME centres/extents must be bit-exact
    nth = A.get_num_tiles_hit_2d(centers, extents, H, W, bw)
    nth_o = oracle.get_num_tiles_hit_2d(to_np(centers), to_np(extents), H, W, bw)
    np.testing.assert_array_equal(to_np(nth), nth_o)


@pytest.mark.parametrize("n", [0, 1, 5, 2047, 2048, 2049, 100000, 1 << 20])
def test_cumsum_exact(n):
    g = torch.Generator().manual_seed(n)
    v = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32).to(DEV)
    total, cum = U.compute_cumulative_intersects(v)
    want = np.cumsum(to_np(v), dtype=np.int32)
    np.testing.assert_array_equal(to_np(cum), want)
    assert total == (int(want[-1]) if n else 0)


@pytest.mark.parametrize("m", [1, 2, 31, 33, 4095, 4096, 4097, 50000, 300001])
@pytest.mark.parametrize("kind", ["random64", "tile_depth", "ties", "negative", "constant"])
def test_sort_pairs_exact_and_stable(m, kind):
    rng = np.random.default_rng(m * 7 + len(kind))
    if kind == "random64":
        keys = rng.integers(-(1 << 62), 1 << 62, m, dtype=np.int64)
    elif kind == "tile_depth":
        tiles = rng.integers(0, 8160, m, dtype=np.int64)
        depth = rng.uniform(6, 10, m).astype(np.float32).view(np.int32).astype(np.int64)
        keys = (tiles << 32) | depth
    elif kind == "ties":
        keys = rng.integers(0, 7, m, dtype=np.int64) << 32 | rng.integers(0, 3, m, dtype=np.int64)
    elif kind == "negative":
        # negative depths sign-extend over the tile bits (reference forward.cu:48): full signed order required
        tiles = rng.integers(0, 100, m, dtype=np.int64)
        depth = rng.uniform(-3, 3, m).astype(np.float32).view(np.int32).astype(np.int64)
        keys = (tiles << 32) | depth
    else:
        keys = np.full(m, 123456789012345, dtype=np.int64)
    vals = np.arange(m, dtype=np.int32)
    ks, vs = U.sort_pairs(torch.from_numpy(keys).to(DEV), torch.from_numpy(vals).to(DEV))
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(to_np(ks), keys[order])
    np.testing.assert_array_equal(to_np(vs), vals[order])
    # and against torch's own CUDA sort (what the reference calls, utils.py:159-160)
    ts, perm = torch.sort(torch.from_numpy(keys).to(DEV), stable=True)
    assert torch.equal(ts, ks) and torch.equal(torch.from_numpy(vals).to(DEV)[perm], vs)


def test_sort_end_bit_and_device_count():
    from gstex_cuda_b200 import _lib
    rng = np.random.default_rng(3)
    m = 100000
    keys = (rng.integers(0, 8160, m, dtype=np.int64) << 32) | rng.uniform(0.5, 20, m).astype(np.float32).view(np.int32)
    vals = rng.integers(0, 1 << 20, m).astype(np.int32)
    k, v = torch.from_numpy(keys).to(DEV), torch.from_numpy(vals).to(DEV)
    ks, vs = U.sort_pairs(k, v, end_bit=32 + 13)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(to_np(ks), keys[order])
    np.testing.assert_array_equal(to_np(vs), vals[order])
    # device-side count: only the first `cnt` elements take part, the tail of the output is untouched
    cnt = 61234
    lib = _lib.load()
    ko, vo = torch.full_like(k, -7), torch.full_like(v, -7)
    temp = torch.empty((lib.gstex_sort_temp_bytes(m),), dtype=torch.uint8, device=DEV)
    dcnt = torch.tensor([cnt], dtype=torch.int32, device=DEV)
    rc = lib.gstex_sort_pairs(m, k.data_ptr(), v.data_ptr(), ko.data_ptr(), vo.data_ptr(), 64, dcnt.data_ptr(),
                              temp.data_ptr(), temp.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    order = np.argsort(keys[:cnt], kind="stable")
    np.testing.assert_array_equal(to_np(ko)[:cnt], keys[:cnt][order])
    np.testing.assert_array_equal(to_np(vo)[:cnt], vals[:cnt][order])
    assert (to_np(ko)[cnt:] == -7).all() and (to_np(vo)[cnt:] == -7).all()


@pytest.mark.parametrize("n,W,H,bw", [(10, 32, 32, 16), (300, 96, 160, 16), (2000, 250, 130, 16), (500, 100, 60, 8),
                                      (20000, 512, 512, 16)])
def test_binning_pipeline_bit_exact(n, W, H, bw):
    s = random_small_scene(n, W, H, seed=n + 1, device=DEV)
    b = bin_cuda(s, bw)
    # oracle binning on the CUDA centres / extents / depths => every integer output must match exactly
    c, e, d = to_np(b["centers"]), to_np(b["extents"]), to_np(b["depths"])
    nth_o = oracle.get_num_tiles_hit_2d(c, e, H, W, bw)
    m_o, cum_o = oracle.compute_cumulative_intersects(nth_o)
    i_o, g_o, is_o, gs_o, bins_o = oracle.bin_and_sort_gaussians(n, m_o, c, e, d, cum_o, b["tile_bounds"], bw)
    assert b["num_intersects"] == m_o
    for k, want in (("num_tiles_hit", nth_o), ("cum_tiles_hit", cum_o), ("isect_ids", i_o), ("gaussian_ids", g_o),
                    ("isect_ids_sorted", is_o), ("gaussian_ids_sorted", gs_o), ("tile_bins", bins_o)):
        np.testing.assert_array_equal(to_np(b[k]), want, err_msg=k)
    # size-independent properties: keys sorted, ranges partition the list, every id inside its tile's AABB range
    ks = to_np(b["isect_ids_sorted"])
    assert np.all(ks[1:] >= ks[:-1])
    bins = to_np(b["tile_bins"])
    assert int((bins[:, 1] - bins[:, 0]).sum()) == m_o


@pytest.mark.parametrize("n,W,H,bw", [(10, 32, 32, 16), (300, 96, 160, 16), (2000, 250, 130, 16), (500, 100, 60, 8)])
def test_wrapped_binning_bit_exact(n, W, H, bw):
    """wrapped=True (torus tile boxes, forward.cu:34-36, 53-62): keys / ids / tile ranges vs oracle.
    Gaussians are spread past the image border so that boxes really wrap around."""
    s = random_small_scene(n, W, H, seed=n + 3, device=DEV, spread=24.0)
    b = bin_cuda(s, bw)
    c, e, d = to_np(b["centers"]), to_np(b["extents"]), to_np(b["depths"])
    nth = oracle.num_tiles_hit_wrapped(c, e, bw)
    m, cum = oracle.compute_cumulative_intersects(nth)
    tb = b["tile_bounds"]
    i_o, g_o, is_o, gs_o, bins_o = oracle.bin_and_sort_gaussians(n, m, c, e, d, cum, tb, bw, wrapped=True)
    cum_t = torch.from_numpy(cum).to(DEV)
    i_c, g_c, is_c, gs_c, bins_c = U.bin_and_sort_gaussians(n, m, b["centers"], b["extents"], b["depths"], cum_t, tb, bw,
                                                            wrapped=True)
    for name, got, want in (("isect_ids", i_c, i_o), ("gaussian_ids", g_c, g_o), ("isect_ids_sorted", is_c, is_o),
                            ("gaussian_ids_sorted", gs_c, gs_o), ("tile_bins", bins_c, bins_o)):
        np.testing.assert_array_equal(to_np(got), want, err_msg=name)
    tiles = to_np(i_c) >> 32
    assert tiles.min() >= 0 and tiles.max() < tb[0] * tb[1]
    assert m > b["num_intersects"]  # the torus boxes are a superset of the clamped ones (and some do wrap)


def test_fused_project_aabb_count_matches_separate_calls():
    from gstex_cuda_b200 import _lib
    s = synthetic_scene(50000, 640, 360, seed=5, device=DEV)
    n, H, W, bw = s["num_points"], s["H"], s["W"], 16
    fx, fy, cx, cy = s["intrins"]
    centers = torch.empty((n, 2), device=DEV); extents = torch.empty((n, 2), device=DEV)
    depths = torch.empty((n,), device=DEV); nth = torch.empty((n,), dtype=torch.int32, device=DEV)
    rc = _lib.load().gstex_project_aabb_count(n, s["means"].data_ptr(), s["scales"].data_ptr(), 1.0, s["quats"].data_ptr(),
                                              s["viewmat"].data_ptr(), fx, fy, cx, cy, H, W, bw, centers.data_ptr(),
                                              extents.data_ptr(), depths.data_ptr(), nth.data_ptr(), 0,
                                              torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    c2, e2 = A.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], s["intrins"])
    _, d2 = A.project_points(s["means"], s["viewmat"], s["intrins"])
    assert torch.equal(centers, c2) and torch.equal(extents, e2) and torch.equal(depths, d2)
    nth2 = A.get_num_tiles_hit_2d(c2, e2, H, W, bw)  # floor rule == truncation rule after the clamp (bw = 16)
    assert torch.equal(nth, nth2)


def test_backend_error_behaviour():
    """CHECK_INPUT semantics (bindings.h:10-15): CPU or non-contiguous inputs raise RuntimeError."""
    s = random_small_scene(10, 32, 32, device=DEV)
    with pytest.raises(RuntimeError):
        _C.get_aabb_2d(s["means"].cpu(), s["scales"], 1.0, s["quats"], s["viewmat"], *s["intrins"])
    with pytest.raises(RuntimeError):
        _C.get_aabb_2d(s["means"].T.contiguous().T, s["scales"], 1.0, s["quats"], s["viewmat"], *s["intrins"])
    with pytest.raises(RuntimeError):  # dtype is checked where the reference's data_ptr<float>() would throw
        _C.get_aabb_2d(s["means"].double(), s["scales"], 1.0, s["quats"], s["viewmat"], *s["intrins"])


# ---- fused tile binning (csrc/binning_tiles.cu) must reproduce the staged path bit for bit -------------------------
@pytest.mark.parametrize("n,W,H,bw,spread", [
    (0, 64, 64, 16, 16.0), (1, 64, 64, 16, 16.0), (300, 96, 160, 16, 16.0), (5000, 256, 192, 16, 16.0),
    (800, 100, 60, 8, 16.0),
    (6000, 32, 32, 16, 4.0),      # ~1.5-6 k entries per tile: the 64 KB shared-memory class
    (40000, 32, 16, 16, 4.0),     # > 8192 entries per tile: the global-memory class
    (200000, 640, 360, 16, 16.0),
])
def test_bin_tiles_matches_staged_path(n, W, H, bw, spread):
    if n == 0:
        z = torch.zeros((0, 2), device=DEV)
        ids, bins, cnt, _ = U.bin_tiles(z, z, torch.zeros((0,), device=DEV), ((W + bw - 1) // bw, (H + bw - 1) // bw, 1), bw, 16)
        assert int(cnt.item()) == 0 and int(bins.abs().sum().item()) == 0
        return
    s = random_small_scene(n, W, H, seed=n + 5, spread=spread, device=DEV)
    if n >= 300:  # equal depths inside a tile: ties must come out in Gaussian order
        s["means"][::7, 2] = s["means"][3, 2]
    b = bin_cuda(s, bw)
    m = b["num_intersects"]
    ids, bins, cnt, isect = U.bin_tiles(b["centers"], b["extents"], b["depths"], b["tile_bounds"], bw, m + 13,
                                        want_isect_ids=True)
    torch.cuda.synchronize()
    assert int(cnt.item()) == m
    assert torch.equal(bins, b["tile_bins"])
    assert torch.equal(ids[:m], b["gaussian_ids_sorted"])
    assert torch.equal(isect[:m], b["isect_ids_sorted"])
    print(f"  M = {m}, longest tile list = {int((bins[:, 1] - bins[:, 0]).max().item())}")


def test_bin_tiles_capacity_overflow_is_clipped():
    s = random_small_scene(3000, 128, 128, seed=4, device=DEV)
    b = bin_cuda(s, 16)
    m = b["num_intersects"]
    cap = m // 2
    ids, bins, cnt, _ = U.bin_tiles(b["centers"], b["extents"], b["depths"], b["tile_bounds"], 16, cap)
    torch.cuda.synchronize()
    assert int(cnt.item()) == m                     # the true count is reported (overflow is detectable) ...
    assert int(bins.max().item()) <= cap            # ... and nothing points past the buffers
    full = b["tile_bins"]
    whole = (full[:, 1] <= cap) & (full[:, 1] > full[:, 0])   # tiles entirely inside the capacity are intact
    assert torch.equal(bins[whole], full[whole])
    last = int(full[whole][:, 1].max().item())
    assert torch.equal(ids[:last], b["gaussian_ids_sorted"][:last])
