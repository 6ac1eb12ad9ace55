#!/usr/bin/env python
"""Inference benchmark (BASELINE.json configs[4]): 1920x1080 novel-view rendering and 512^3 marching-cubes SDF extraction of a
model trained on the DTU-shaped synthetic scene.  One process per GPU (torchrun for N > 1): image rows / lattice z-slabs are
dealt across ranks (neus2_b200.parallel), no collective on the data path, results gathered at the end.

  python tools/bench_infer.py [--train-steps 1000] [--width 1920 --height 1080] [--mc 512] [--frames 10] [--json out.json]
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
    ap.add_argument("--train-steps", type=int, default=1000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--mc", type=int, default=512)
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--views", type=int, default=49)
    ap.add_argument("--train-width", type=int, default=1600)
    ap.add_argument("--train-height", type=int, default=1200)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import neus2_b200 as N
    from neus2_b200 import parallel, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    data = synthetic.dtu_like_views(n_views=args.views, width=args.train_width, height=args.train_height, seed=0)
    ctx = N.Context(N.default_config(), device=local)       # every rank trains the same model (deterministic up to atomics order) ...
    ctx.init_params()
    ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    for _ in range(args.train_steps):
        ctx.train_step()
    torch.cuda.synchronize()
    if world > 1:                                             # ... and rank 0's weights are broadcast so that all ranks render the same scene
        for t in ctx.buffers()[:3]:
            dist.broadcast(t, 0)
        grid = ctx.debug_buffer("density_grid", "<f4", (128 ** 3,))
        dist.broadcast(grid, 0)
        ctx.set_density_grid(grid.cpu().numpy(), ema_step=ctx.get_state()["density_grid_ema_step"])

    # ---- novel views on an orbit between the training cameras
    W, H = args.width, args.height
    cams = synthetic.camera_ring(args.frames, radius=2.0, elevation_deg=(25.0, 45.0), seed=11)
    focal = 2890.0 * H / 1200.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = rays = 0
    render_ms = []
    for i in range(args.frames + 2):
        rp = ctx.render_params(W, H, camera_matrix=cams[i % args.frames], focal_length=(focal, focal), min_transmittance=0.01, snap_to_pixel_centers=1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        img = parallel.render(ctx, rp, spp=1, rank=rank, world_size=world)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            render_ms.append(float(t.item()))
            if world == 1:
                st = ctx.render_stats()
                samples += st["samples"]; rays += st["rays_alive"]
    coverage = float((img[..., 3] > 0.5).float().mean().item())

    # ---- SDF lattice + marching cubes
    res = (args.mc,) * 3
    lat_ms, mc_ms = [], []
    for i in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        lat = parallel.sdf_lattice(ctx, res, rank=rank, world_size=world)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lat_ms.append(float(t.item()))
        e0.record()
        verts, faces = ctx.marching_cubes(lat)
        e1.record()
        torch.cuda.synchronize()
        mc_ms.append(e0.elapsed_time(e1))
    if rank == 0:
        n_pts = args.mc ** 3
        out = {
            "workload": f"{W}x{H} novel-view render + {args.mc}^3 SDF lattice + marching cubes, model trained {args.train_steps} steps on the DTU-shaped synthetic scene",
            "n_gpus": world, "render_ms_per_frame": float(np.mean(render_ms)), "render_fps": 1e3 / float(np.mean(render_ms)),
            "render_Mpixels_per_s": W * H / float(np.mean(render_ms)) / 1e3, "object_coverage": coverage,
            "lattice_ms": float(np.min(lat_ms)), "lattice_Gpoints_per_s": n_pts / float(np.min(lat_ms)) / 1e6,
            "lattice_alg_GBps": n_pts * (14 * 8 * 4 + 4) / float(np.min(lat_ms)) / 1e6,
            "marching_cubes_ms": float(np.min(mc_ms)), "n_verts": int(len(verts)), "n_faces": int(len(faces)),
        }
        if world == 1 and render_ms:
            out["render_samples_per_frame"] = samples / len(render_ms)
            out["render_rays_alive_per_frame"] = rays / len(render_ms)
            out["render_Msamples_per_s"] = samples / len(render_ms) / float(np.mean(render_ms)) / 1e3
        print(json.dumps(out))
        if args.json:
            with open(args.json, "w") as f:
                json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
