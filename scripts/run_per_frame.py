#!/usr/bin/env python
"""Per-frame reconstruction of a sequence, one frame per GPU at a time (the reference's scripts/run_per_frame.py trains every
`<frame>.json` of a directory independently, one after another; frames are independent units, so here they are dealt across
the GPUs of the box: rank r takes frames r, r + W, r + 2W, ... -- replicas only, no collective).

  python scripts/run_per_frame.py --base_dir data/sequence --output_dir out [--config base.json] [--gpus 8] [--n_steps 2000]
      [--marching_cubes_res 256] [--synthetic 20]     # --synthetic N: write N frames of a moving analytic scene into base_dir first

Per frame it writes <output_dir>/<frame>/{snapshot.msgpack, mesh.obj, stats.json}; rank 0 prints one JSON line with frames/min.
"""
import argparse
import glob
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_synthetic_sequence(base_dir, n_frames, views=16, res=256):
    """BASELINE.json configs[3]: the analytic two-sphere scene translated by 0.01*k and rotated 2 deg * k about z, same cameras."""
    import numpy as np
    from neus2_b200 import loader, synthetic
    os.makedirs(base_dir, exist_ok=True)
    for k in range(n_frames):
        a = np.deg2rad(2.0 * k)
        c0 = np.array([0.5 + 0.01 * k, 0.5, 0.5])
        off = np.array([0.18 * np.cos(a) - 0.05 * np.sin(a), 0.18 * np.sin(a) + 0.05 * np.cos(a), 0.1])
        data = synthetic.render_sphere_views(n_views=views, width=res, height=res, seed=0, elevation_deg=(-55.0, 65.0), center=(0.5, 0.5, 0.5),
                                             spheres=[(c0, 0.25), (c0 + off, 0.25 * 0.45)])
        d = os.path.join(base_dir, f"frame_{k:04d}")
        loader.write_dataset(d, data, scale=0.5, offset=(0.5, 0.5, 0.5), from_na=True, name="transforms.json")
        js = json.load(open(os.path.join(d, "transforms.json")))
        for fr in js["frames"]:
            fr["file_path"] = f"frame_{k:04d}/" + fr["file_path"]
        json.dump(js, open(os.path.join(base_dir, f"{k:04d}.json"), "w"))


def worker(rank, world, args, frames):
    import torch
    torch.cuda.set_device(rank)
    from neus2_b200.testbed import Testbed, TestbedMode
    from neus2_b200 import loader
    cfg = loader.load_json(args.config) if args.config and os.path.exists(args.config) else None
    done = []
    for scene in frames[rank::world]:
        num = os.path.basename(scene).split(".")[0]
        out = os.path.join(args.output_dir, num)
        os.makedirs(out, exist_ok=True)
        t0 = time.time()
        tb = Testbed(TestbedMode.Nerf, device=rank)
        tb.load_training_data(scene)
        if cfg is not None:
            tb.reload_network_from_json(cfg)
        else:
            sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
            from run import DEFAULT_NETWORK
            tb.reload_network_from_json(DEFAULT_NETWORK)
        tb.first_frame_max_training_step = args.n_steps
        t1 = time.time()
        while tb.training_step < args.n_steps:
            tb.frame()
        torch.cuda.synchronize()
        t2 = time.time()
        tb.save_snapshot(os.path.join(out, "snapshot.msgpack"), False)
        n_tri = tb.compute_and_save_marching_cubes_mesh(os.path.join(out, "mesh.obj"), [args.marching_cubes_res] * 3)
        t3 = time.time()
        st = dict(frame=num, gpu=rank, load_s=t1 - t0, train_s=t2 - t1, export_s=t3 - t2, loss=tb.loss, triangles=n_tri)
        json.dump(st, open(os.path.join(out, "stats.json"), "w"))
        done.append(st)
        tb._ctx.close()
    return done


def _entry(rank, world, args, frames, q):
    q.put((rank, worker(rank, world, args, frames)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--base_dir", default=".")
    ap.add_argument("--output_dir", default="neus2_exp")
    ap.add_argument("--config", default="")
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--n_steps", type=int, default=2000)
    ap.add_argument("--marching_cubes_res", type=int, default=256)
    ap.add_argument("--synthetic", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.multiprocessing as mp
    if args.synthetic:
        write_synthetic_sequence(args.base_dir, args.synthetic)
    frames = sorted(glob.glob(os.path.join(args.base_dir, "*.json")))
    if not frames:
        raise SystemExit(f"no *.json frames under {args.base_dir}")
    world = args.gpus or torch.cuda.device_count()
    world = max(1, min(world, len(frames)))
    t0 = time.time()
    if world == 1:
        results = worker(0, 1, args, frames)
    else:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_entry, args=(r, world, args, frames, q)) for r in range(world)]
        for p in procs:
            p.start()
        results = []
        for _ in procs:
            results += q.get()[1]
        for p in procs:
            p.join()
    wall = time.time() - t0
    print(json.dumps({"frames": len(results), "n_gpus": world, "wall_s": wall, "frames_per_min": 60.0 * len(results) / wall,
                      "mean_train_s": sum(r["train_s"] for r in results) / len(results), "mean_load_s": sum(r["load_s"] for r in results) / len(results),
                      "mean_export_s": sum(r["export_s"] for r in results) / len(results), "n_steps": args.n_steps}))


if __name__ == "__main__":
    main()
