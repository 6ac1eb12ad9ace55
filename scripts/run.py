#!/usr/bin/env python
"""Command-line driver with the flags of the reference's scripts/run.py (the subset that applies to the NeuS path):

  python scripts/run.py --scene data/scene/transforms.json [--network configs/base.json] --n_steps 2000 \
      [--save_snapshot out.msgpack] [--load_snapshot in.msgpack] [--save_mesh mesh.obj --marching_cubes_res 512] \
      [--screenshot_dir shots --screenshot_frames 0 1 2 --screenshot_spp 1 --width W --height H] [--test]

It drives neus2_b200.testbed.Testbed exactly as the reference script drives pyngp.Testbed (load_training_data,
reload_network_from_file, frame() loop with tqdm-style progress, save_snapshot, compute_and_save_marching_cubes_mesh,
set_camera_to_training_view + render).  `--test` prints PSNR of every training view against its ground truth.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

DEFAULT_NETWORK = {
    "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "nested": {
        "otype": "Adam", "learning_rate": 1e-3, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
    "encoding": {"otype": "HashGrid", "n_levels": 14, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16,
                 "top_resolution": 2048, "valid_level_scale": 0.02, "base_valid_level_scale": 0.2, "base_training_step": 100},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1},
    "rgb_network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
    "hyperparams": {"first_frame_max_training_step": 2000, "mask_loss_weight": 0.0, "ek_loss_weight": 0.01},
}


def parse_args():
    p = argparse.ArgumentParser(description="Train / render / mesh a NeuS2 scene with neus2_b200")
    p.add_argument("--scene", "--training_data", default="", help="transforms.json or a directory of them")
    p.add_argument("--network", default="", help="network config json (configs/nerf/base.json schema); default = base.json values")
    p.add_argument("--load_snapshot", default="")
    p.add_argument("--save_snapshot", default="")
    p.add_argument("--n_steps", type=int, default=-1)
    p.add_argument("--save_mesh", default="")
    p.add_argument("--marching_cubes_res", type=int, default=256)
    p.add_argument("--screenshot_dir", default="")
    p.add_argument("--screenshot_frames", nargs="*", type=int, default=[])
    p.add_argument("--screenshot_spp", type=int, default=1)
    p.add_argument("--width", type=int, default=0)
    p.add_argument("--height", type=int, default=0)
    p.add_argument("--test", action="store_true")
    p.add_argument("--mask_loss_weight", type=float, default=None)
    return p.parse_args()


def main():
    args = parse_args()
    from neus2_b200.testbed import Testbed, TestbedMode
    testbed = Testbed(TestbedMode.Nerf)
    if args.scene:
        testbed.load_training_data(args.scene)
    if args.load_snapshot:
        testbed.load_snapshot(args.load_snapshot)
    elif args.network:
        testbed.reload_network_from_file(args.network)
    else:
        cfg = dict(DEFAULT_NETWORK)
        if args.mask_loss_weight is not None:
            cfg["hyperparams"] = dict(cfg["hyperparams"], mask_loss_weight=args.mask_loss_weight)
        testbed.reload_network_from_json(cfg)
    testbed.shall_train = True
    n_steps = args.n_steps if args.n_steps >= 0 else (0 if args.load_snapshot else testbed.first_frame_max_training_step)
    testbed.first_frame_max_training_step = max(testbed.first_frame_max_training_step, n_steps)
    t0 = time.time()
    last = t0
    while testbed.training_step < n_steps and testbed.frame():
        now = time.time()
        if now - last > 1.0 or testbed.training_step == n_steps:
            print(f"step {testbed.training_step}/{n_steps}  loss {testbed.loss:.6f}  ek {testbed.ek_loss:.6f}  {now - t0:.1f}s", flush=True)
            last = now
    if args.save_snapshot:
        testbed.save_snapshot(args.save_snapshot, False)
        print("saved snapshot", args.save_snapshot)
    if args.save_mesh:
        res = args.marching_cubes_res
        n = testbed.compute_and_save_marching_cubes_mesh(args.save_mesh, [res, res, res])
        print(f"saved mesh {args.save_mesh}: {n} triangles")
    d = testbed._data
    if args.test and d is not None:
        testbed.snap_to_pixel_centers = True
        testbed.nerf.rendering_min_transmittance = 1e-4
        testbed.background_color = [0.0, 0.0, 0.0, 1.0]
        psnrs = []
        for i in range(len(d["paths"])):
            testbed.set_camera_to_training_view(i)
            w, h = d["resolution"]
            img = testbed.render(w, h, 1, False)
            gt = d["pixels"][i].astype(np.float32) / 255.0
            mse = float(np.mean((np.clip(img[..., :3], 0, 1) - gt[..., :3] * gt[..., 3:4]) ** 2))
            psnrs.append(-10.0 * np.log10(max(mse, 1e-12)))
        print(f"PSNR over {len(psnrs)} training views: mean {np.mean(psnrs):.2f} dB, min {np.min(psnrs):.2f} dB")
    if args.screenshot_dir and d is not None:
        from PIL import Image
        os.makedirs(args.screenshot_dir, exist_ok=True)
        testbed.snap_to_pixel_centers = True
        for i in (args.screenshot_frames or range(len(d["paths"]))):
            testbed.set_camera_to_training_view(i)
            w = args.width or d["resolution"][0]
            h = args.height or d["resolution"][1]
            img = testbed.render(w, h, args.screenshot_spp, False)
            out = os.path.join(args.screenshot_dir, os.path.splitext(os.path.basename(d["paths"][i]))[0] + ".png")
            Image.fromarray((np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8), "RGBA").save(out)
            print("wrote", out)


if __name__ == "__main__":
    main()
