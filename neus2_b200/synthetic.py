"""Synthetic multi-view RGBA data rendered from an analytic SDF (there is no network for real datasets).

Produces what the reference's loader (src/nerf_loader.cu) would hand the Testbed after `from_na` conversion
(nerf_loader.h:112-120): RGBA8 images (straight sRGB colour + alpha = hit mask, as read_rgba expects,
common_device.cuh:635-656), per-image focal length in pixels, relative principal point, and camera-to-world
3x4 matrices in ngp coordinates (x right, y down, z forward; the scene lives in the unit cube around 0.5).
"""
import numpy as np


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    if np.linalg.norm(right) < 1e-6:
        right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    m = np.zeros((3, 4), dtype=np.float64)
    m[:, 0], m[:, 1], m[:, 2], m[:, 3] = right, down, fwd, eye
    return m


def camera_ring(n_views, radius=1.3, center=(0.5, 0.5, 0.5), elevation_deg=(15.0, 50.0), seed=0):
    rs = np.random.RandomState(seed)
    xf = []
    for i in range(n_views):
        az = 2 * np.pi * i / n_views
        el = np.deg2rad(elevation_deg[0] + (elevation_deg[1] - elevation_deg[0]) * rs.rand())
        eye = np.asarray(center) + radius * np.array([np.cos(az) * np.cos(el), np.sin(az) * np.cos(el), np.sin(el)])
        xf.append(look_at(eye, center))
    return np.stack(xf).astype(np.float32)


def render_sphere_views(n_views=16, width=256, height=256, focal=None, sphere_radius=0.25, center=(0.5, 0.5, 0.5),
                        cam_radius=1.3, seed=0, second_sphere=True, principal=(0.5, 0.5), elevation_deg=(15.0, 50.0), spheres=None):
    """Returns dict(pixels uint8 [n,h,w,4], focal [n,2], principal [n,2], xforms [n,3,4])."""
    if focal is None:
        focal = 1.1 * width * cam_radius / (4 * sphere_radius) / 1.3
    fx, fy = (focal, focal) if np.isscalar(focal) else focal
    xforms = camera_ring(n_views, radius=cam_radius, center=center, elevation_deg=elevation_deg, seed=seed)
    cx, cy = principal
    ys, xs = np.meshgrid((np.arange(height) + 0.5) / height, (np.arange(width) + 0.5) / width, indexing="ij")
    d_cam = np.stack([(xs - cx) * width / fx, (ys - cy) * height / fy, np.ones_like(xs)], axis=-1)
    if spheres is None:
        spheres = [(np.asarray(center, dtype=np.float64), sphere_radius)]
        if second_sphere:
            spheres.append((np.asarray(center, dtype=np.float64) + np.array([0.18, 0.05, 0.1]), sphere_radius * 0.45))
    else:
        spheres = [(np.asarray(c, dtype=np.float64), float(r)) for c, r in spheres]
    pixels = np.zeros((n_views, height, width, 4), dtype=np.uint8)
    light = np.array([0.3, 0.5, 0.8]) / np.linalg.norm([0.3, 0.5, 0.8])
    for v in range(n_views):
        R, o = xforms[v, :, :3].astype(np.float64), xforms[v, :, 3].astype(np.float64)
        d = d_cam @ R.T
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        best_t = np.full((height, width), np.inf)
        normal = np.zeros((height, width, 3))
        for c, r in spheres:
            oc = o - c
            b = (d * oc).sum(-1)
            disc = b * b - (oc @ oc - r * r)
            t = -b - np.sqrt(np.maximum(disc, 0))
            hit = (disc > 0) & (t > 0) & (t < best_t)
            p = o + t[..., None] * d
            n = (p - c) / r
            best_t = np.where(hit, t, best_t)
            normal = np.where(hit[..., None], n, normal)
        hit = np.isfinite(best_t)
        albedo = 0.5 + 0.5 * normal
        shade = 0.25 + 0.75 * np.clip((normal * light).sum(-1), 0, 1)
        rgb_lin = np.clip(albedo * shade[..., None], 0, 1)
        srgb = np.where(rgb_lin < 0.0031308, 12.92 * rgb_lin, 1.055 * np.power(np.maximum(rgb_lin, 1e-12), 1 / 2.4) - 0.055)
        pixels[v, ..., :3] = np.where(hit[..., None], np.round(srgb * 255), 0).astype(np.uint8)
        pixels[v, ..., 3] = np.where(hit, 255, 0).astype(np.uint8)
    return dict(pixels=pixels, focal=np.tile(np.array([[fx, fy]], dtype=np.float32), (n_views, 1)),
                principal=np.tile(np.array([[cx, cy]], dtype=np.float32), (n_views, 1)), xforms=xforms, spheres=spheres)


def dtu_like_views(n_views=49, width=1600, height=1200, seed=0):
    """DTU-Scan24-shaped data (BASELINE.json config 2): 49 views, 1600x1200, focal ~2890 px."""
    f = (2892.33 * width / 1600.0, 2883.18 * height / 1200.0)
    return render_sphere_views(n_views=n_views, width=width, height=height, focal=f, sphere_radius=0.25, cam_radius=2.2, seed=seed,
                               principal=(823.2 / 1600.0, 619.07 / 1200.0))


def sphere_sdf(points, spheres):
    d = np.full(points.shape[0], np.inf)
    for c, r in spheres:
        d = np.minimum(d, np.linalg.norm(points - c, axis=-1) - r)
    return d


# ---- a second analytic scene with flat faces and sharp edges: an axis-aligned box with a sphere on top -------------------------
def box_sdf(points, center, half):
    q = np.abs(points - np.asarray(center)) - np.asarray(half)
    return np.linalg.norm(np.maximum(q, 0.0), axis=-1) + np.minimum(q.max(axis=-1), 0.0)


def box_scene_sdf(points, scene):
    d = box_sdf(points, scene["box_center"], scene["box_half"])
    for c, r in scene["spheres"]:
        d = np.minimum(d, np.linalg.norm(points - c, axis=-1) - r)
    return d


def render_box_views(n_views=16, width=256, height=256, focal=None, cam_radius=1.3, seed=0, elevation_deg=(-55.0, 65.0)):
    """Like render_sphere_views, for a box (0.34 x 0.26 x 0.2) with a small sphere sitting on it; exact ray casting."""
    center = np.array([0.5, 0.5, 0.5])
    scene = dict(box_center=center - np.array([0.0, 0.0, 0.05]), box_half=np.array([0.17, 0.13, 0.10]),
                 spheres=[(center + np.array([0.05, -0.02, 0.11]), 0.09)])
    if focal is None:
        focal = 1.1 * width * cam_radius / 1.0 / 1.3
    xforms = camera_ring(n_views, radius=cam_radius, center=tuple(center), elevation_deg=elevation_deg, seed=seed)
    ys, xs = np.meshgrid((np.arange(height) + 0.5) / height, (np.arange(width) + 0.5) / width, indexing="ij")
    d_cam = np.stack([(xs - 0.5) * width / focal, (ys - 0.5) * height / focal, np.ones_like(xs)], axis=-1)
    pixels = np.zeros((n_views, height, width, 4), dtype=np.uint8)
    light = np.array([0.3, 0.5, 0.8]) / np.linalg.norm([0.3, 0.5, 0.8])
    bmin, bmax = scene["box_center"] - scene["box_half"], scene["box_center"] + scene["box_half"]
    for v in range(n_views):
        R, o = xforms[v, :, :3].astype(np.float64), xforms[v, :, 3].astype(np.float64)
        d = d_cam @ R.T
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0, t1 = (bmin - o) / d, (bmax - o) / d
        tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
        t_in, t_out = tn.max(-1), tf.min(-1)
        hit_box = (t_in < t_out) & (t_in > 0)
        axis = tn.argmax(-1)
        n_box = np.zeros(d.shape)
        np.put_along_axis(n_box, axis[..., None], -np.sign(np.take_along_axis(d, axis[..., None], -1)), -1)
        best_t = np.where(hit_box, t_in, np.inf)
        normal = np.where(hit_box[..., None], n_box, 0.0)
        for c, r in scene["spheres"]:
            oc = o - c
            b = (d * oc).sum(-1)
            disc = b * b - (oc @ oc - r * r)
            t = -b - np.sqrt(np.maximum(disc, 0))
            hit = (disc > 0) & (t > 0) & (t < best_t)
            p = o + t[..., None] * d
            best_t = np.where(hit, t, best_t)
            normal = np.where(hit[..., None], (p - c) / r, normal)
        hit = np.isfinite(best_t)
        albedo = 0.5 + 0.5 * normal
        shade = 0.25 + 0.75 * np.clip((normal * light).sum(-1), 0, 1)
        rgb_lin = np.clip(albedo * shade[..., None], 0, 1)
        srgb = np.where(rgb_lin < 0.0031308, 12.92 * rgb_lin, 1.055 * np.power(np.maximum(rgb_lin, 1e-12), 1 / 2.4) - 0.055)
        pixels[v, ..., :3] = np.where(hit[..., None], np.round(srgb * 255), 0).astype(np.uint8)
        pixels[v, ..., 3] = np.where(hit, 255, 0).astype(np.uint8)
    return dict(pixels=pixels, focal=np.full((n_views, 2), focal, dtype=np.float32), principal=np.full((n_views, 2), 0.5, dtype=np.float32),
                xforms=xforms, scene=scene)
