"""Generates tests/golden/ngp_helpers_reference.npz from the REFERENCE's own device helper headers
(include/neural-graphics-primitives/common_device.cuh: srgb_to_linear, linear_to_srgb, image_pos, read_rgba, pixel_to_ray,
fov_to_focal_length, focal_length_to_fov; bounding_box.cuh: BoundingBox::ray_intersect, contains; nerf_loader.h: the NerfDataset matrix
conventions), compiled as host code from where they
lie into oracle/_ref/libref_ngp_host.so by oracle/Makefile.ref.  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_ngp_helpers_golden.py tests/golden/ngp_helpers_reference.npz
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_ngp_host.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def main(out_path):
    L = C.CDLL(LIB)
    for f in ("refn_srgb_to_linear", "refn_linear_to_srgb", "refn_fov_to_focal_length", "refn_focal_length_to_fov"):
        getattr(L, f).restype = C.c_float
    L.refn_srgb_to_linear.argtypes = [C.c_float]; L.refn_linear_to_srgb.argtypes = [C.c_float]
    L.refn_fov_to_focal_length.argtypes = [C.c_int, C.c_float]; L.refn_focal_length_to_fov.argtypes = [C.c_int, C.c_float]
    rs = np.random.RandomState(99)
    out = {}
    # colour transfer functions: every 8-bit level (what read_rgba feeds them) + random values around the knees
    x = np.concatenate([np.arange(256, dtype=np.float32) * np.float32(1.0 / 255.0), rs.uniform(0, 1.2, 512).astype(np.float32),
                        np.array([0.04045, 0.0404501, 0.0031308, 0.00313081, 0.0, 1.0], np.float32)])
    out["srgb_x"] = x
    out["srgb_to_linear"] = np.array([L.refn_srgb_to_linear(float(v)) for v in x], np.float32)
    out["linear_to_srgb"] = np.array([L.refn_linear_to_srgb(float(v)) for v in x], np.float32)
    res = np.array([64, 800, 1200, 1600, 1920], np.int32)
    deg = rs.uniform(10, 120, len(res)).astype(np.float32)
    out["fov_res"], out["fov_deg"] = res, deg
    out["fov_to_focal"] = np.array([L.refn_fov_to_focal_length(int(r), float(d)) for r, d in zip(res, deg)], np.float32)
    out["focal_to_fov"] = np.array([L.refn_focal_length_to_fov(int(r), float(f)) for r, f in zip(res, out["fov_to_focal"])], np.float32)
    # AABB tests: rays from outside and inside, axis-parallel directions (division by +-0), points on the faces
    n = 1024
    aabb = np.array([0, 0, 0, 1, 1, 1], np.float32)
    o = rs.uniform(-1.5, 2.5, (n, 3)).astype(np.float32)
    d = rs.normal(0, 1, (n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:32, 0] = 0.0; d[32:48, 1] = -0.0; d[48:64, :2] = 0.0
    o[64:96] = rs.uniform(0, 1, (32, 3))
    t = np.zeros((n, 2), np.float32)
    for i in range(n):
        L.refn_ray_intersect(_p(aabb), _p(o[i]), _p(d[i]), _p(t[i]))
    pts = rs.uniform(-0.1, 1.1, (n, 3)).astype(np.float32)
    pts[:8] = [[0, 0.5, 0.5], [1, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 1, 0.5], [0.5, 0.5, 0], [0.5, 0.5, 1], [1, 1, 1], [0, 0, 0]]
    inside = np.array([L.refn_contains(_p(aabb), _p(pts[i])) for i in range(n)], np.int32)
    out.update(aabb=aabb, ray_o=o, ray_d=d, ray_t=t, points=pts, inside=inside)
    # images: RGBA8 incl. the hot-pink mask word 0x00FF00FF and zero alpha
    w, h, n_img = 37, 29, 3
    px = rs.randint(0, 256, (n_img, h, w, 4)).astype(np.uint8)
    px[0, 0, 0] = [255, 0, 255, 0]; px[1, 5, 7] = [255, 0, 255, 0]; px[2, 3, 3, 3] = 0
    xy = rs.uniform(-0.05, 1.05, (256, 2)).astype(np.float32)
    xy[:4] = [[0, 0], [1, 1], [0.999999, 0.5], [0.5, 1.0]]
    ip = np.zeros((len(xy), 2), np.int32)
    for i in range(len(xy)):
        L.refn_image_pos(_p(xy[i]), w, h, _p(ip[i]))
    rgba = np.zeros((n_img, h, w, 4), np.float32)
    tmp = np.zeros(4, np.float32)
    for im in range(n_img):
        for yy in range(h):
            for xx in range(w):
                L.refn_read_rgba(xx, yy, w, h, _p(px), C.c_uint32(im), _p(tmp))
                rgba[im, yy, xx] = tmp
    out.update(img_pixels=px, img_xy=xy, img_pos=ip, img_rgba=rgba)
    # pixel_to_ray: render cameras (jittered, spp > 0) and training-style use (snap, spp = 0)
    cams = []
    for k in range(6):
        R = np.linalg.qr(rs.normal(0, 1, (3, 3)))[0].astype(np.float32)
        cam = np.concatenate([R, rs.uniform(-1, 2, (3, 1)).astype(np.float32)], axis=1).astype(np.float32)
        W, H = int(rs.randint(16, 2000)), int(rs.randint(16, 1400))
        focal = rs.uniform(0.5, 3.0, 2).astype(np.float32) * np.float32(W)
        sc = rs.uniform(0.3, 0.7, 2).astype(np.float32)
        spp = int(rs.randint(0, 1000)); snap = int(k % 2)
        pix = np.stack([rs.randint(0, W, 64), rs.randint(0, H, 64)], 1).astype(np.int32)
        rays = np.zeros((64, 6), np.float32)
        for i in range(64):
            L.refn_pixel_to_ray(C.c_uint32(spp), int(pix[i, 0]), int(pix[i, 1]), W, H, _p(focal), _p(np.ascontiguousarray(cam.reshape(12))), _p(sc), snap, _p(rays[i]))
        cams.append(dict(cam=cam, wh=np.array([W, H], np.int32), focal=focal, sc=sc, spp=np.array(spp), snap=np.array(snap), pix=pix, rays=rays))
    for k, c in enumerate(cams):
        for name, v in c.items():
            out[f"cam{k}_{name}"] = v
    out["n_cams"] = np.array(len(cams))
    # dataset coordinate conventions (NerfDataset::nerf_matrix_to_ngp / ngp_matrix_to_nerf, nerf_loader.h:112-158): default axis cycle, from_na, from_mitsuba
    LL = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_loader_host.so"))
    LL.refl_nerf_scale.restype = C.c_float
    out["nerf_scale"] = np.float32(LL.refl_nerf_scale())
    mats = rs.normal(size=(9, 3, 4)).astype(np.float32)
    offs = rs.uniform(0, 1, (9, 3)).astype(np.float32)
    scs = rs.uniform(0.2, 2, 9).astype(np.float32)
    fwd, inv = np.zeros((9, 12), np.float32), np.zeros((9, 12), np.float32)
    for i in range(9):
        LL.refl_matrix(_p(np.ascontiguousarray(mats[i].reshape(12))), C.c_float(float(scs[i])), _p(np.ascontiguousarray(offs[i])), i % 3, 0, _p(fwd[i]))
        LL.refl_matrix(_p(np.ascontiguousarray(mats[i].reshape(12))), C.c_float(float(scs[i])), _p(np.ascontiguousarray(offs[i])), i % 3, 1, _p(inv[i]))
    out.update(loader_mats=mats, loader_offsets=offs, loader_scales=scs, loader_to_ngp=fwd.reshape(9, 3, 4), loader_to_nerf=inv.reshape(9, 3, 4))
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ngp_helpers_reference.npz"))
