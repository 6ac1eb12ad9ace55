"""transforms.json loader -- the data format on the input side of the training path.

Restates `load_nerf` (src/nerf_loader.cu:196-745) and `NerfDataset::nerf_matrix_to_ngp`
(include/neural-graphics-primitives/nerf_loader.h:112-135) for the NeuS2 data convention (README.md:77-148):
8-bit RGBA images, per-frame `transform_matrix` (camera-to-world) and optional 4x4 `intrinsic_matrix`, global
`from_na`, `scale`, `offset`, `aabb_scale`, `w`, `h`, `fl_x` / `fl_y` / `camera_angle_x|y`, `n_frames` (`cx` / `cy` are ignored, as in the reference),
`white_transparent`, `black_transparent`, `<file>.alpha.<ext>` side images and `dynamic_mask_<file>.png` masks.
Images are decoded by a thread pool (PIL releases the GIL) into ONE pinned host buffer so that the upload to the GPU
is a single asynchronous copy.  Everything is float32 arithmetic in the reference's order of operations.
"""
import json
import math
import os
import re
from concurrent.futures import ThreadPoolExecutor

import numpy as np

NERF_SCALE = np.float32(0.33)   # nerf_loader.h:31


def _strip_json_comments(text):
    """nlohmann::json::parse(..., ignore_comments=true): drop // and /* */ comments outside strings."""
    out, i, n, in_str = [], 0, len(text), False
    while i < n:
        c = text[i]
        if in_str:
            out.append(c)
            if c == "\\" and i + 1 < n:
                out.append(text[i + 1]); i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True; out.append(c)
        elif c == "/" and i + 1 < n and text[i + 1] == "/":
            while i < n and text[i] != "\n":
                i += 1
            continue
        elif c == "/" and i + 1 < n and text[i + 1] == "*":
            i = text.find("*/", i + 2)
            i = n if i < 0 else i + 2
            continue
        else:
            out.append(c)
        i += 1
    return "".join(out)


def load_json(path):
    with open(path) as f:
        return json.loads(_strip_json_comments(f.read()))


def nerf_matrix_to_ngp(m, scale, offset, from_na=False, from_mitsuba=False):
    """nerf_loader.h:112-135 on a 3x4 float32 matrix."""
    r = np.array(m, dtype=np.float32)[:3, :4].copy()
    r[:, 1] *= np.float32(-1)
    r[:, 2] *= np.float32(-1)
    r[:, 3] = r[:, 3] * np.float32(scale) + np.asarray(offset, dtype=np.float32)
    if from_na:
        r[:, 1] *= np.float32(-1)
        r[:, 2] *= np.float32(-1)
    elif from_mitsuba:
        r[:, 0] *= np.float32(-1)
        r[:, 2] *= np.float32(-1)
    else:   # cycle axes xyz <- yzx
        r = r[[1, 2, 0], :]
    return np.ascontiguousarray(r)


def ngp_matrix_to_nerf(m, scale, offset, from_na=False, from_mitsuba=False):
    """nerf_loader.h:137-158."""
    r = np.array(m, dtype=np.float32)[:3, :4].copy()
    if from_na:
        r[:, 1] *= np.float32(-1)
        r[:, 2] *= np.float32(-1)
    elif from_mitsuba:
        r[:, 0] *= np.float32(-1)
        r[:, 2] *= np.float32(-1)
    else:
        r = r[[2, 0, 1], :]
    r[:, 1] *= np.float32(-1)
    r[:, 2] *= np.float32(-1)
    r[:, 3] = (r[:, 3] - np.asarray(offset, dtype=np.float32)) / np.float32(scale)
    return np.ascontiguousarray(r)


def fov_to_focal_length(resolution, degrees):
    """common.h: 0.5 * resolution / tan(0.5 * degrees * pi / 180) in float."""
    return np.float32(0.5) * np.float32(resolution) / np.float32(math.tan(np.float32(0.5) * np.float32(degrees) * np.float32(math.pi) / np.float32(180.0)))


def _srgb_to_linear(s):
    return s / 12.92 if s <= 0.04045 else ((s + 0.055) / 1.055) ** 2.4


def _decode(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGBA"), dtype=np.uint8)


def _resolve_image(basepath, file_path):
    p = os.path.join(basepath, file_path.replace("\\", "/"))
    if os.path.splitext(p)[1] == "":
        if os.path.exists(p + ".png"):
            return p + ".png"
        raise RuntimeError("Could not find image file: " + p + ".png")
    if not os.path.exists(p):
        raise RuntimeError("image not found: " + p)
    return p


def list_json_paths(path):
    """Testbed::load_training_data accepts a .json file or a directory whose *.json files are all loaded (testbed.cu:196-228)."""
    if os.path.isdir(path):
        files = sorted(os.path.join(path, f) for f in os.listdir(path) if f.endswith(".json"))
        if not files:
            raise RuntimeError(f"Data path '{path}' does not contain any transforms json")
        return files
    if not os.path.exists(path):
        raise RuntimeError(f"Data path '{path}' does not exist.")
    return [path]


def load_nerf(path, n_threads=None, pin=True):
    """Returns a dict: pixels uint8 [n,h,w,4] (pinned torch-backed when `pin` and torch are available), focal [n,2] (pixels),
    principal [n,2] (relative), xforms [n,3,4] (ngp convention), plus scale, offset, aabb_scale, from_na, paths, resolution."""
    jsonpaths = list_json_paths(path)
    jobs, focal, principal, xforms, names = [], [], [], [], []
    scale, offset, aabb_scale, from_na = None, None, 1, False
    for jp in jsonpaths:
        js = load_json(jp)
        if not isinstance(js.get("frames"), list):
            continue
        base = os.path.dirname(os.path.abspath(jp))
        frames = js["frames"]
        if "n_frames" in js:
            frames = frames[: min(len(frames), int(js["n_frames"]))]
        if "from_na" in js:            # presence, not value (nerf_loader.cu:346-348)
            from_na = True
        sc = np.float32(js["scale"]) if "scale" in js else (NERF_SCALE if scale is None else scale)
        if "offset" in js:
            o = js["offset"]
            off = np.array(o if isinstance(o, list) else [o, o, o], dtype=np.float32)
        else:
            off = np.array([0.5, 0.5, 0.5], dtype=np.float32) if offset is None else offset
        scale, offset = sc, off
        if "aabb_scale" in js:
            aabb_scale = int(js["aabb_scale"])
        # The reference parses "cx" / "cy" (nerf_loader.cu:431-437) but never stores them: its assignment of the principal point to the image
        # metadata is commented out (:703), so the centre stays at (0.5, 0.5) unless a frame carries an "intrinsic_matrix" (:678-686).
        # Same here -- checked against the reference's own loader (tests/test_golden_loader.py).
        pp0 = np.array([0.5, 0.5], dtype=np.float32)
        for fr in frames:
            img_path = _resolve_image(base, fr["file_path"])
            root, ext = os.path.splitext(img_path)
            alpha = os.path.join(base, fr["file_path"] + ".alpha" + ext)
            mask = os.path.join(os.path.dirname(img_path), "dynamic_mask_" + os.path.basename(root) + ".png")
            jobs.append((img_path, alpha if os.path.exists(alpha) else None, mask if os.path.exists(mask) else None,
                         bool(js.get("white_transparent", False)), bool(js.get("black_transparent", False))))
            names.append(img_path)
            m = fr.get("transform_matrix_start", fr.get("transform_matrix"))
            xforms.append(nerf_matrix_to_ngp(np.array(m, dtype=np.float32), sc, off, from_na))
            w_h = (js.get("w"), js.get("h"))

            def read_fl(axis, res):
                if axis + "_fov" in fr:
                    return fov_to_focal_length(res, fr[axis + "_fov"])
                if "fl_" + axis in js:
                    return np.float32(js["fl_" + axis])
                if "camera_angle_" + axis in js:
                    return fov_to_focal_length(res, np.float32(js["camera_angle_" + axis]) * np.float32(180) / np.float32(math.pi))
                return np.float32(0)
            focal.append((fr, js, read_fl, w_h))
            principal.append(pp0.copy())
    if not jobs:
        raise RuntimeError("no frames found in " + str(path))

    def work(job):
        img_path, alpha, mask, white_t, black_t = job
        img = _decode(img_path).copy()
        if alpha is not None:
            a = _decode(alpha)
            if a.shape != img.shape:
                raise RuntimeError("Alpha image has wrong resolution: " + alpha)
            lut = np.array([np.uint8(255.0 * _srgb_to_linear(v / 255.0)) for v in range(256)], dtype=np.uint8)
            img[..., 3] = lut[a[..., 0]]
        if mask is not None:
            mk = _decode(mask)
            if mk.shape != img.shape:
                raise RuntimeError("Mask image has wrong resolution: " + mask)
            img[mk[..., 0] != 0] = (0xFF, 0x00, 0xFF, 0x00)       # hot pink, convert_rgba32 :75-78
        if white_t:
            img[(img[..., :3] == 255).all(-1), 3] = 0
        if black_t:
            img[(img[..., :3] == 0).all(-1), 3] = 0
        return img

    first = work(jobs[0])
    h, w = first.shape[:2]
    n = len(jobs)
    pixels = None
    if pin:
        try:
            import torch
            pixels = torch.empty((n, h, w, 4), dtype=torch.uint8, pin_memory=torch.cuda.is_available()).numpy()
        except Exception:
            pixels = None
    if pixels is None:
        pixels = np.empty((n, h, w, 4), dtype=np.uint8)
    pixels[0] = first

    def fill(i):
        img = work(jobs[i])
        if img.shape[:2] != (h, w):
            raise RuntimeError(f"{names[i]}: this build trains on images of one resolution ({w}x{h}), got {img.shape[1]}x{img.shape[0]}")
        pixels[i] = img
    with ThreadPoolExecutor(max_workers=n_threads or min(32, os.cpu_count() or 4)) as pool:
        list(pool.map(fill, range(1, n)))

    focal_out = np.zeros((n, 2), dtype=np.float32)
    principal_out = np.stack(principal).astype(np.float32)
    for i, (fr, js, read_fl, _) in enumerate(focal):
        x_fl, y_fl = read_fl("x", w), read_fl("y", h)
        if x_fl != 0:
            focal_out[i] = (x_fl, y_fl if y_fl != 0 else x_fl)
        elif y_fl != 0:
            focal_out[i] = (y_fl, y_fl)
        elif "intrinsic_matrix" in fr:
            K = fr["intrinsic_matrix"]
            focal_out[i] = (np.float32(K[0][0]), np.float32(K[1][1]))
            principal_out[i] = (np.float32(K[0][2]) / np.float32(js["w"]), np.float32(K[1][2]) / np.float32(js["h"]))
        else:
            raise RuntimeError("Couldn't read fov.")
    return dict(pixels=pixels, focal=focal_out, principal=principal_out, xforms=np.stack(xforms).astype(np.float32), scale=float(scale),
                offset=np.asarray(offset, dtype=np.float32), aabb_scale=aabb_scale, from_na=from_na, paths=names, resolution=(w, h))


def write_dataset(out_dir, data, scale=0.5, offset=(0.5, 0.5, 0.5), from_na=True, name="transforms.json"):
    """Writes a synthetic scene (neus2_b200.synthetic) in the NeuS2 data convention: PNGs + transforms.json with per-frame
    `transform_matrix` (the inverse of nerf_matrix_to_ngp) and 4x4 `intrinsic_matrix`."""
    from PIL import Image
    os.makedirs(os.path.join(out_dir, "images"), exist_ok=True)
    n, h, w, _ = data["pixels"].shape
    frames = []
    for i in range(n):
        fn = f"images/{i:06d}.png"
        Image.fromarray(data["pixels"][i], "RGBA").save(os.path.join(out_dir, fn))
        m = ngp_matrix_to_nerf(data["xforms"][i], scale, offset, from_na).astype(np.float64)
        K = [[float(data["focal"][i, 0]), 0.0, float(data["principal"][i, 0] * np.float32(w)), 0.0],
             [0.0, float(data["focal"][i, 1]), float(data["principal"][i, 1] * np.float32(h)), 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]
        frames.append({"file_path": fn, "transform_matrix": m.tolist() + [[0.0, 0.0, 0.0, 1.0]], "intrinsic_matrix": K})
    js = {"w": w, "h": h, "aabb_scale": 1, "scale": float(scale), "offset": [float(v) for v in offset], "frames": frames}
    if from_na:
        js["from_na"] = True
    with open(os.path.join(out_dir, name), "w") as f:
        json.dump(js, f)
    return os.path.join(out_dir, name)


def config_from_json(cfg_json):
    """Keys of configs/nerf/*.json this build consumes (src/testbed.cu:2084-2199) -> keyword arguments of neus2_b200.default_config."""
    kw = {}
    enc = cfg_json.get("encoding", {})
    for src, dst in (("n_levels", "n_levels"), ("log2_hashmap_size", "log2_hashmap_size"), ("base_resolution", "base_resolution"),
                     ("valid_level_scale", "valid_level_scale"), ("base_valid_level_scale", "base_valid_level_scale"), ("base_training_step", "base_training_step")):
        if src in enc:
            kw[dst] = type(enc[src])(enc[src])
    if "top_resolution" in enc:
        kw["top_resolution"] = float(enc["top_resolution"])
    if "per_level_scale" in enc:
        kw["per_level_scale"] = float(enc["per_level_scale"])
    if enc.get("otype", "HashGrid") not in ("HashGrid", "Grid"):
        raise RuntimeError(f"encoding otype {enc.get('otype')} is not built (hash grid only)")
    if enc.get("n_features_per_level", 2) != 2:
        raise RuntimeError("n_features_per_level must be 2")
    for name in ("network", "rgb_network"):
        net = cfg_json.get(name, {})
        if net.get("n_neurons", 64) != 64 or net.get("otype", "FullyFusedMLP") != "FullyFusedMLP" or net.get("activation", "ReLU") != "ReLU":
            raise RuntimeError(f"{name}: this build fixes 64-wide ReLU FullyFusedMLPs")
    if cfg_json.get("network", {}).get("n_hidden_layers", 1) != 1 or cfg_json.get("rgb_network", {}).get("n_hidden_layers", 2) != 2:
        raise RuntimeError("n_hidden_layers must be 1 (sdf) / 2 (rgb)")
    opt = cfg_json.get("optimizer", {})
    if "decay" in opt:
        kw["ema_decay"] = float(opt["decay"])
    while True:
        if opt.get("otype", "") == "ExponentialDecay":       # defaults of exponential_decay.h:145-148
            kw["decay_start"] = int(opt.get("decay_start", 10000))
            kw["decay_interval"] = int(opt.get("decay_interval", 10000))
            kw["decay_end"] = int(opt.get("decay_end", 10000000))
            kw["decay_base"] = float(opt.get("decay_base", 0.1))
        if "nested" not in opt:
            break
        opt = opt["nested"]
    for src, dst in (("learning_rate", "learning_rate"), ("beta1", "beta1"), ("beta2", "beta2"), ("epsilon", "epsilon"), ("l2_reg", "l2_reg")):
        if src in opt:
            kw[dst] = float(opt[src])
    # the DeltaNetwork's own trainer: config["globalmove"]["optimizer"], else the main optimizer config (testbed.cu:2320-2323)
    gopt = cfg_json.get("globalmove", {}).get("optimizer", cfg_json.get("optimizer", {}))
    while True:
        if gopt.get("otype", "") == "ExponentialDecay":
            kw["delta_decay_start"] = int(gopt.get("decay_start", 10000))
            kw["delta_decay_interval"] = int(gopt.get("decay_interval", 10000))
            kw["delta_decay_base"] = float(gopt.get("decay_base", 0.1))
        if "nested" not in gopt:
            break
        gopt = gopt["nested"]
    for src, dst in (("learning_rate", "delta_learning_rate"), ("beta1", "delta_beta1"), ("beta2", "delta_beta2"), ("epsilon", "delta_epsilon")):
        if src in gopt:
            kw[dst] = float(gopt[src])
    hp = cfg_json.get("hyperparams", {})
    for src, dst in (("mask_loss_weight", "mask_loss_weight"), ("ek_loss_weight", "ek_loss_weight")):
        if src in hp:
            kw[dst] = float(hp[src])
    if "anneal_end" in hp:
        kw["anneal_end"] = int(hp["anneal_end"])
    if hp:                                                  # src/testbed.cu:2123: the CODE default is true; configs/nerf/base.json sets false
        kw["reset_density_grid_after_global_movement"] = int(bool(hp.get("reset_density_grid_after_global_movement", True)))
    return kw
