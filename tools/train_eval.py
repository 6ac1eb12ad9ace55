#!/usr/bin/env python
"""End-to-end acceptance run (BASELINE.json configs[0]): train the fixed 2000-step schedule on the synthetic analytic-SDF
scene (16 views 256x256), then render a training view with the EMA weights (PSNR against the ground-truth image) and
extract the zero level set with marching cubes (Chamfer distance to the analytic surface).

  python tools/train_eval.py [--steps 2000] [--views 16] [--res 256] [--mc 256] [--json out.json]

The reference cannot be run here for a side-by-side Chamfer (its ngp layer does not compile offline), so the
numbers are absolute: distance of the extracted surface to the analytic one, in units of the scene cube edge.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--views", type=int, default=16)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--mc", type=int, default=256)
    ap.add_argument("--json", default=None)
    ap.add_argument("--box", action="store_true", help="a box with a sphere on top instead of the two spheres: flat faces, sharp edges, a concave junction")
    ap.add_argument("--dtu", action="store_true", help="BASELINE.json configs[1]: the DTU-Scan24-shaped scene (49 views 1600x1200, cameras on a cap above the object)")
    ap.add_argument("--mask-weight", type=float, default=0.0)
    ap.add_argument("--full-forward", action="store_true", help="evaluate every marched sample in the first pass (reference behaviour) instead of the early cut")
    args = ap.parse_args()

    import torch
    from scipy.spatial import cKDTree
    import neus2_b200 as N
    from neus2_b200 import synthetic

    if args.dtu:
        data = synthetic.dtu_like_views(n_views=49, width=1600, height=1200, seed=0)
        args.views = 49
    elif args.box:
        data = synthetic.render_box_views(n_views=args.views, width=args.res, height=args.res, seed=0)
    else:
        data = synthetic.render_sphere_views(n_views=args.views, width=args.res, height=args.res, seed=0, elevation_deg=(-55.0, 65.0))
    H, W = data["pixels"].shape[1:3]
    true_sdf = (lambda pts: synthetic.box_scene_sdf(pts, data["scene"])) if args.box else (lambda pts: synthetic.sphere_sdf(pts, data["spheres"]))
    ctx = N.Context(N.default_config(mask_loss_weight=args.mask_weight, forward_early_termination=0 if args.full_forward else 1))
    ctx.init_params()
    ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    torch.cuda.synchronize()
    log = []
    t0 = time.perf_counter()
    rays = 0
    for i in range(args.steps):
        st = ctx.train_step()
        rays += st.rays_per_batch
        if st.loss_valid and (i % 160 == 0 or i == args.steps - 1):
            log.append(dict(step=int(st.training_step), loss=float(st.loss), ek_loss=float(st.ek_loss), rays=int(st.rays_per_batch),
                            samples=int(st.measured_batch_size_before_compaction), compacted=int(st.measured_batch_size)))
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0

    # ---- render every training view with the EMA weights; PSNR against the ground truth composited on black
    psnr = []
    t0 = time.perf_counter()
    for v in range(args.views):
        rp = ctx.training_view_params(data["xforms"][v], data["focal"][v], data["principal"][v], (W, H), min_transmittance=1e-4,
                                      background_color=(0, 0, 0, 1), to_srgb=1)
        img = ctx.render(rp, spp=1)
        gt = torch.from_numpy(data["pixels"][v]).cuda().float() / 255.0
        gt_rgb = gt[..., :3] * gt[..., 3:4]     # straight sRGB colour x alpha over black (sRGB compositing as in the eval scripts)
        mse = ((img[..., :3].clamp(0, 1) - gt_rgb) ** 2).mean().item()
        psnr.append(-10.0 * np.log10(max(mse, 1e-12)))
    torch.cuda.synchronize()
    t_render = time.perf_counter() - t0
    rstats = ctx.render_stats()

    # ---- mesh: SDF lattice (EMA weights) + marching cubes, Chamfer against the analytic surface
    t0 = time.perf_counter()
    lat = ctx.sdf_lattice((args.mc,) * 3)
    torch.cuda.synchronize()
    t_lattice = time.perf_counter() - t0
    t0 = time.perf_counter()
    verts, faces = ctx.marching_cubes((-lat).contiguous(), thresh=0.0)     # inside = sdf < 0
    torch.cuda.synchronize()
    t_mc = time.perf_counter() - t0
    V = verts.cpu().numpy().astype(np.float64)
    if args.dtu and len(V):
        # the cameras sit on a cap above the object: the underside is never observed, judge the surface they see (z above the lowest camera ray)
        seen = V[:, 2] > 0.5 - 0.05
        V = V[seen]
    d_mesh_to_gt = np.abs(true_sdf(V))
    # analytic surface samples
    rs = np.random.RandomState(0)
    if args.box:
        # rejection: project random points near the surface along the numerical SDF gradient
        P = rs.uniform(0.2, 0.8, (400000, 3))
        P = P[np.abs(true_sdf(P)) < 0.02]
        for _ in range(3):
            e = 1e-4
            g = np.stack([(true_sdf(P + np.eye(3)[k] * e) - true_sdf(P - np.eye(3)[k] * e)) / (2 * e) for k in range(3)], -1)
            P = P - true_sdf(P)[:, None] * g / np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-9)
        P = P[np.abs(true_sdf(P)) < 1e-5]
    else:
        pts = []
        for c, r in data["spheres"]:
            n = int(40000 * (r / data["spheres"][0][1]) ** 2)
            d = rs.normal(size=(n, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            pts.append(c + r * d)
        P = np.concatenate(pts)
        P = P[np.abs(true_sdf(P)) < 1e-9]
        if args.dtu:
            P = P[P[:, 2] > 0.5 - 0.05]
    d_gt_to_mesh = cKDTree(V).query(P)[0] if len(V) else np.full(len(P), np.inf)
    chamfer = 0.5 * (d_mesh_to_gt.mean() + d_gt_to_mesh.mean())

    out = dict(steps=args.steps, views=args.views, res=[int(W), int(H)], scene="dtu-shaped" if args.dtu else ("box + sphere" if args.box else "two spheres, cameras all around"), train_s=t_train, ms_per_step=1e3 * t_train / args.steps, rays_per_s=rays / t_train,
               psnr_mean=float(np.mean(psnr)), psnr_min=float(np.min(psnr)), render_s_per_view=t_render / args.views,
               render_last_view=rstats, mc_res=args.mc, lattice_s=t_lattice, mc_s=t_mc, n_verts=int(len(V)), n_faces=int(len(faces)),
               mesh_to_gt_mean=float(d_mesh_to_gt.mean()), mesh_to_gt_p99=float(np.percentile(d_mesh_to_gt, 99)) if len(V) else None,
               gt_to_mesh_mean=float(d_gt_to_mesh.mean()), chamfer=float(chamfer), voxel=1.0 / args.mc, log=log)
    print(json.dumps(out))
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
