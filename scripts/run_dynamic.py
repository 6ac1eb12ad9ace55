#!/usr/bin/env python
"""Incremental reconstruction of a dynamic sequence (the reference's scripts/run_dynamic.py, :236-345): `--scene` is a directory
with one transforms json per time frame.  Frame 0 trains for first_frame_max_training_step steps; every later frame continues
from the previous frame's inference weights for next_frame_max_training_step steps and -- with hyperparams.predict_global_movement --
first registers the frame against the canonical model with a rigid motion (DeltaNetwork alone for
predict_global_movement_training_step steps, then jointly; the motions accumulate from frame to frame).

  python scripts/run_dynamic.py --name seq --scene data/sequence [--network base.json] [--save_snapshot_per_frame] [--save_mesh]
      [--marching_cubes_res 256] [--test_psnr] [--synthetic 5]     # --synthetic N: write N frames of a moving analytic scene first

At the end of every frame it writes, like the reference, under output/<name>/:
  checkpoints/frame_<k>.msgpack (--save_snapshot_per_frame), pred_transform/frame_<k>.txt (Testbed::save_transform),
  mesh/frame_<k:04>.obj (--save_mesh), and a line in training_log.txt / eval_log.txt; the last stdout line is a JSON summary.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

DYNAMIC_HYPERPARAMS = {"first_frame_max_training_step": 2000, "next_frame_max_training_step": 2000, "predict_global_movement": True,
                       "predict_global_movement_training_step": 50, "finetune_global_movement": True, "mask_loss_weight": 0.0, "ek_loss_weight": 0.01}
GLOBALMOVE = {"optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 50, "decay_interval": 25, "decay_base": 0.99,
              "nested": {"otype": "Adam", "learning_rate": 1e-4, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-10, "l2_reg": 1e-10}}}}


def parse_args():
    p = argparse.ArgumentParser(description="Train a dynamic sequence frame by frame with neus2_b200")
    p.add_argument("--name", default="neus", type=str)
    p.add_argument("--scene", "--training_data", default="", help="directory with one transforms json per time frame")
    p.add_argument("--network", default="", help="network config json (configs/nerf/base.json schema); default = base.json values")
    p.add_argument("--output_path", default="", help="default: output/<name>")
    p.add_argument("--save_snapshot_per_frame", action="store_true")
    p.add_argument("--save_mesh", action="store_true")
    p.add_argument("--marching_cubes_res", type=int, default=256)
    p.add_argument("--test_psnr", action="store_true", help="PSNR of every frame's training views after the frame has trained")
    p.add_argument("--first_frame_steps", type=int, default=-1)
    p.add_argument("--next_frame_steps", type=int, default=-1)
    p.add_argument("--no_global_movement", action="store_true", help="hyperparams.predict_global_movement = false")
    p.add_argument("--synthetic", type=int, default=0)
    p.add_argument("--synthetic_res", type=int, default=256)
    p.add_argument("--synthetic_views", type=int, default=16)
    return p.parse_args()


def psnr_training_views(testbed, max_views=8):
    """cal_psnr of the reference's script, on (up to) max_views of the frame's own training views, black background."""
    import numpy as np
    d = testbed._data
    testbed.background_color = [0.0, 0.0, 0.0, 1.0]
    testbed.snap_to_pixel_centers = True
    h, w = int(d["resolution"][1]), int(d["resolution"][0])
    vals = []
    for v in range(0, len(d["xforms"]), max(1, len(d["xforms"]) // max_views)):
        testbed.set_camera_to_training_view(v)
        img = np.clip(testbed.render(w, h, 1, False)[..., :3], 0, 1)
        gt = d["pixels"][v].astype(np.float32) / 255.0
        mse = float(np.mean((img - gt[..., :3] * gt[..., 3:4]) ** 2))
        vals.append(-10.0 * np.log10(max(mse, 1e-12)))
    return float(np.mean(vals))


def main():
    args = parse_args()
    import torch
    from neus2_b200 import loader
    from neus2_b200.testbed import Testbed, TestbedMode
    from run import DEFAULT_NETWORK
    import run_per_frame
    if args.synthetic:
        run_per_frame.write_synthetic_sequence(args.scene, args.synthetic, views=args.synthetic_views, res=args.synthetic_res)
    out = args.output_path or os.path.join("output", args.name)
    for sub in ("checkpoints", "mesh", "pred_transform"):
        os.makedirs(os.path.join(out, sub), exist_ok=True)
    testbed = Testbed(TestbedMode.Nerf)
    testbed.load_training_data(args.scene)
    if args.network:
        cfg = loader.load_json(args.network)
    else:
        cfg = dict(DEFAULT_NETWORK, globalmove=GLOBALMOVE, hyperparams=dict(DYNAMIC_HYPERPARAMS))
        cfg["optimizer"] = json.loads(json.dumps(cfg["optimizer"]))
        cfg["optimizer"]["nested"]["nested"]["after_learning_rate"] = 1e-3
    hp = dict(cfg.get("hyperparams", {}))
    if args.first_frame_steps >= 0:
        hp["first_frame_max_training_step"] = args.first_frame_steps
    if args.next_frame_steps >= 0:
        hp["next_frame_max_training_step"] = args.next_frame_steps
    if args.no_global_movement:
        hp["predict_global_movement"] = False
    cfg["hyperparams"] = hp
    testbed.reload_network_from_json(cfg)
    testbed.shall_train = True
    train_log = open(os.path.join(out, "training_log.txt"), "w")
    eval_log = open(os.path.join(out, "eval_log.txt"), "w")
    print("begin training!", file=train_log, flush=True)
    frames = []
    t_all = now = time.monotonic()
    total_steps = 0
    while testbed.frame():
        total_steps += 1
        k = testbed.current_training_time_frame
        budget = testbed.first_frame_max_training_step if k == 0 else testbed.next_frame_max_training_step
        if total_steps % 200 == 0:
            print(f"training time_frame: {k}, training step: {testbed.training_step}, training_time: {time.monotonic() - now:.2f}, trainig_rgb_loss:{testbed.loss:.6f}", flush=True)
        if testbed.training_step == budget:
            torch.cuda.synchronize()
            train_s = time.monotonic() - now
            print(f"training frame {k} for time :{train_s}\n", file=train_log, flush=True)
            rec = dict(frame=k, train_s=train_s, loss=testbed.loss)
            if args.save_snapshot_per_frame:
                testbed.save_snapshot(os.path.join(out, "checkpoints", f"frame_{k}.msgpack"), False)
            testbed.save_transform(os.path.join(out, "pred_transform", f"frame_{k}.txt"))
            R, t = testbed.global_transform()
            rec["rotation"], rec["transition"] = R.tolist(), t.tolist()
            if args.save_mesh:
                res = args.marching_cubes_res
                rec["triangles"] = testbed.compute_and_save_marching_cubes_mesh(os.path.join(out, "mesh", f"frame_{k:04}.obj"), [res, res, res])
            if args.test_psnr:
                rec["psnr"] = psnr_training_views(testbed)
                print(f"frame {k}: PSNR {rec['psnr']:.2f}", file=eval_log, flush=True)
            frames.append(rec)
            print(f"frame {k} done in {train_s:.2f}s loss {testbed.loss:.6f}" + (f" psnr {rec['psnr']:.2f}" if "psnr" in rec else ""), flush=True)
            now = time.monotonic()
            if k == testbed.all_training_time_frame - 1:
                break
    wall = time.monotonic() - t_all
    later = [f["train_s"] for f in frames[1:]]
    summary = dict(frames=len(frames), wall_s=wall, first_frame_train_s=frames[0]["train_s"] if frames else None,
                   mean_next_frame_train_s=(sum(later) / len(later)) if later else None,
                   predict_global_movement=bool(testbed.predict_global_movement), per_frame=frames)
    json.dump(summary, open(os.path.join(out, "summary.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "per_frame"} | ({"mean_psnr": sum(f["psnr"] for f in frames) / len(frames)} if args.test_psnr and frames else {})))


if __name__ == "__main__":
    main()
