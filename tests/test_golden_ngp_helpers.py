"""CPU pin of the oracle's ray / image / colour helpers against golden vectors produced by the REFERENCE's own headers:
tests/golden/ngp_helpers_reference.npz is frozen from include/neural-graphics-primitives/common_device.cuh (srgb_to_linear,
linear_to_srgb, image_pos, read_rgba, pixel_to_ray, fov conversions) and bounding_box.cuh (BoundingBox::ray_intersect / contains)
compiled as host code from where they lie (oracle/Makefile.ref, tests/golden/make_ngp_helpers_golden.py).  Comparisons, integer results
and the scalar float functions are bit-exact; the 3x3 products differ by nvcc's FMA contraction, which the oracle mirrors and the
host build does not (<= 2 ulp of the largest term)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from neus2_b200 import loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngp_helpers_reference.npz")
REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_ngp_host.so")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def test_colour_transfer_functions_bit_exact(gold):
    L = O.lib()
    x = gold["srgb_x"]
    a = np.array([L.orc_srgb_to_linear(C.c_float(float(v))) for v in x], np.float32)
    b = np.array([L.orc_linear_to_srgb(C.c_float(float(v))) for v in x], np.float32)
    assert np.array_equal(_bits(a), _bits(gold["srgb_to_linear"]))
    assert np.array_equal(_bits(b), _bits(gold["linear_to_srgb"]))


def test_fov_conversions(gold):
    f = np.array([loader.fov_to_focal_length(int(r), float(d)) for r, d in zip(gold["fov_res"], gold["fov_deg"])], np.float32)
    assert np.abs(f - gold["fov_to_focal"]).max() <= 2e-6 * np.abs(gold["fov_to_focal"]).max()      # tanf (reference) vs math.tan rounded to float
    back = np.array([2.0 * 180.0 / np.pi * np.arctan(float(r) / (float(ff) * 2.0)) for r, ff in zip(gold["fov_res"], gold["fov_to_focal"])], np.float32)
    assert np.abs(back - gold["focal_to_fov"]).max() < 1e-4 and np.abs(gold["focal_to_fov"] - gold["fov_deg"]).max() < 1e-4


def test_aabb_ray_intersect_and_contains_bit_exact(gold):
    L = O.lib()
    aabb = np.ascontiguousarray(gold["aabb"])
    o, d, want = gold["ray_o"], gold["ray_d"], gold["ray_t"]
    got = np.zeros_like(want)
    with np.errstate(all="ignore"):
        for i in range(len(o)):
            L.orc_ray_intersect(_p(aabb), _p(np.ascontiguousarray(o[i])), _p(np.ascontiguousarray(d[i])), _p(got[i]))
    assert np.array_equal(_bits(got), _bits(want))                       # incl. the +-inf / FLT_MAX cases of axis-parallel rays
    assert (want[:, 0] == np.finfo(np.float32).max).any() and (want[:, 0] < 10).any()
    inside = np.array([L.orc_aabb_contains(_p(aabb), _p(np.ascontiguousarray(p))) for p in gold["points"]], np.int32)
    assert np.array_equal(inside, gold["inside"]) and inside[:8].all()   # faces and corners are inside


def test_image_pos_and_read_rgba_bit_exact(gold):
    L = O.lib()
    px = np.ascontiguousarray(gold["img_pixels"])
    n_img, h, w, _ = px.shape
    ds = O.make_dataset(px, np.ones((n_img, 2), np.float32), np.full((n_img, 2), 0.5, np.float32), np.zeros((n_img, 3, 4), np.float32))
    ip = np.zeros_like(gold["img_pos"])
    for i, xy in enumerate(gold["img_xy"]):
        L.orc_image_pos(_p(np.ascontiguousarray(xy)), w, h, _p(ip[i]))
    assert np.array_equal(ip, gold["img_pos"])
    out = np.zeros((n_img, h, w, 4), np.float32)
    tmp = np.zeros(4, np.float32)
    for im in range(n_img):
        for y in range(h):
            for x in range(w):
                L.orc_read_rgba(C.byref(ds), C.c_uint32(im), x, y, _p(tmp))
                out[im, y, x] = tmp
    assert np.array_equal(_bits(out), _bits(gold["img_rgba"]))
    assert np.array_equal(out[0, 0, 0], [-1, -1, -1, -1]) and np.array_equal(out[2, 3, 3], [0, 0, 0, 0])   # mask pixel, zero alpha


def test_pixel_to_ray_matches_reference(gold):
    L = O.lib()
    for k in range(int(gold["n_cams"])):
        cam, (W, H), focal, sc = gold[f"cam{k}_cam"], gold[f"cam{k}_wh"], gold[f"cam{k}_focal"], gold[f"cam{k}_sc"]
        spp, snap, pix, want = int(gold[f"cam{k}_spp"]), int(gold[f"cam{k}_snap"]), gold[f"cam{k}_pix"], gold[f"cam{k}_rays"]
        rc = O.make_render_cam(int(W), int(H), focal, sc, cam, sample_index=spp, snap=bool(snap))
        got = np.zeros_like(want)
        for i, (x, y) in enumerate(pix):
            L.orc_pixel_to_ray(C.byref(rc), int(x), int(y), _p(got[i]))
        assert np.array_equal(_bits(got[:, :3]), _bits(want[:, :3]))              # origin = camera column 3
        scale = np.abs(want[:, 3:]).max()
        assert np.abs(got[:, 3:] - want[:, 3:]).max() <= 3e-7 * max(scale, 1.0)   # 3x3 product: FMA contraction vs plain IEEE
        # the training path's ray (make_ray: snapped pixel centre, principal point = screen centre) is the same function
        if snap:
            ds = O.make_dataset(np.zeros((1, int(H), int(W), 4), np.uint8), focal[None].copy(), sc[None].copy(), cam[None].copy())
            tr = np.zeros(6, np.float32)
            for i, (x, y) in enumerate(pix[:8]):
                xy = np.array([(x + 0.5) / W, (y + 0.5) / H], np.float32)
                L.orc_make_ray(C.byref(ds), C.c_uint32(0), _p(xy), _p(tr))
                assert np.abs(tr - want[i]).max() <= 3e-6 * max(scale, 1.0)


def test_dataset_matrix_conventions_bit_exact(gold):
    """transforms.json camera matrices -> ngp coordinates and back (NerfDataset::nerf_matrix_to_ngp / ngp_matrix_to_nerf): axis cycle,
    from_na, from_mitsuba -- what neus2_b200.loader applies to every frame of a dataset."""
    assert np.float32(loader.NERF_SCALE) == gold["nerf_scale"]
    for i in range(9):
        kw = [dict(), dict(from_na=True), dict(from_mitsuba=True)][i % 3]
        m, sc, off = gold["loader_mats"][i], gold["loader_scales"][i], gold["loader_offsets"][i]
        a = np.asarray(loader.nerf_matrix_to_ngp(m, sc, off, **kw), np.float32)
        b = np.asarray(loader.ngp_matrix_to_nerf(m, sc, off, **kw), np.float32)
        assert np.array_equal(_bits(a), _bits(gold["loader_to_ngp"][i])), i
        assert np.array_equal(_bits(b), _bits(gold["loader_to_nerf"][i])), i


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference/include")), reason="needs the reference tree (oracle/_ref/libref_ngp_host.so)")
def test_committed_fixture_is_what_the_reference_source_produces(gold, tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ngp_helpers_golden", os.path.join(os.path.dirname(GOLD), "make_ngp_helpers_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(gold.files)
    for name in gold.files:
        assert np.array_equal(fresh[name], gold[name], equal_nan=True), name
