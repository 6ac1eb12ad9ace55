"""Generates tests/golden/loader_reference.npz by running the REFERENCE's own dataset loader (src/nerf_loader.cu load_nerf / set_training_image,
compiled as host code from where it lies: oracle/Makefile.ref -> oracle/_ref/libref_loadnerf_host.so) on small synthetic datasets that exercise
the transforms.json dialects of the NeuS / instant-ngp data format.  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_loader_golden.py

The datasets are rebuilt from seeds by the test (build_case below); stored: intrinsics, camera matrices, scale / offset / aabb_scale / from_na and
the decoded RGBA8 pixels (small images) as the reference's loader left them.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

W, H = 20, 12
CASES = ["neus_from_na_intrinsic_matrix", "nerf_synthetic_camera_angle", "colmap_fl_cx_aabb4_scalar_offset", "alpha_mask_transparency", "n_frames_two_jsons"]


def _pose(rs):
    q, _ = np.linalg.qr(rs.normal(size=(3, 3)))
    m = np.eye(4)
    m[:3, :3] = q * np.sign(np.linalg.det(q))
    m[:3, 3] = rs.uniform(-2, 2, 3)
    return m.tolist()


def _image(rs, path, opaque_fraction=0.6, w=W, h=H):
    from PIL import Image
    px = rs.randint(0, 256, (h, w, 4)).astype(np.uint8)
    px[..., 3] = np.where(rs.rand(h, w) < opaque_fraction, 255, rs.randint(0, 256, (h, w))).astype(np.uint8)
    px[0, :3, :3] = 255; px[1, :3, :3] = 0                         # a few pure white / pure black pixels for the transparency switches
    os.makedirs(os.path.dirname(path), exist_ok=True)
    Image.fromarray(px, "RGBA").save(path)


def build_case(name, root):
    """Writes the dataset `name` under `root`; returns the list of transforms json paths in load order."""
    from PIL import Image
    rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    if name == "neus_from_na_intrinsic_matrix":
        frames = []
        for i in range(3):
            _image(rs, os.path.join(d, "images", f"{i:06d}.png"))
            K = [[300.5 + i, 0, W / 2 + 0.25 * i, 0], [0, 301.25 + i, H / 2 - 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
            frames.append({"file_path": f"images/{i:06d}.png", "transform_matrix": _pose(rs), "intrinsic_matrix": K})
        js = {"from_na": True, "w": W, "h": H, "aabb_scale": 1, "scale": 0.5, "offset": [0.5, 0.5, 0.5], "frames": frames}
        paths = [os.path.join(d, "transforms.json")]
        json.dump(js, open(paths[0], "w"))
    elif name == "nerf_synthetic_camera_angle":
        frames = []
        for i in range(2):
            _image(rs, os.path.join(d, "train", f"r_{i}.png"))
            frames.append({"file_path": f"./train/r_{i}", "rotation": 0.1, "transform_matrix": _pose(rs)})      # no extension: ".png" is appended
        js = {"camera_angle_x": 0.6911112070083618, "frames": frames}
        paths = [os.path.join(d, "transforms_train.json")]
        json.dump(js, open(paths[0], "w"))
    elif name == "colmap_fl_cx_aabb4_scalar_offset":
        frames = []
        for i in range(2):
            _image(rs, os.path.join(d, "images", f"{i:04d}.png"))
            frames.append({"file_path": f"images\\\\{i:04d}.png".replace("\\\\", "\\"), "sharpness_unused": 1.0, "transform_matrix": _pose(rs)})   # Windows separator
        js = {"fl_x": 410.125, "fl_y": 408.5, "cx": 9.75, "cy": 6.5, "w": W, "h": H, "k1": 0.0, "aabb_scale": 4, "scale": 0.25, "offset": 0.4, "frames": frames}
        paths = [os.path.join(d, "transforms.json")]
        json.dump(js, open(paths[0], "w"))
    elif name == "alpha_mask_transparency":
        frames = []
        for i in range(2):
            p = os.path.join(d, "images", f"{i:03d}.png")
            _image(rs, p)
            frames.append({"file_path": f"images/{i:03d}.png", "transform_matrix": _pose(rs)})
        a = rs.randint(0, 256, (H, W, 4)).astype(np.uint8); a[..., 3] = 255
        Image.fromarray(a, "RGBA").save(os.path.join(d, "images", "000.png.alpha.png"))                # <file_path>.alpha.<ext>: sRGB-coded alpha channel
        m = np.zeros((H, W, 4), np.uint8); m[2:5, 4:9, 0] = 200; m[..., 3] = 255
        Image.fromarray(m, "RGBA").save(os.path.join(d, "images", "dynamic_mask_001.png"))             # dynamic_mask_<stem>.png: masked-out pixels
        js = {"fl_x": 200.0, "w": W, "h": H, "white_transparent": True, "black_transparent": True, "frames": frames}
        paths = [os.path.join(d, "transforms.json")]
        json.dump(js, open(paths[0], "w"))
    elif name == "n_frames_two_jsons":
        paths = []
        for part, n_keep in (("a", 2), ("b", 1)):
            frames = []
            for i in range(3):
                _image(rs, os.path.join(d, part, f"{i}.png"))
                frames.append({"file_path": f"{part}/{i}.png", "transform_matrix": _pose(rs), "x_fov": 35.0 + i, "y_fov": 22.0})
            # "from_na" counts by presence.  Same scale in both files: the reference converts a file's matrices in thread-pool tasks that read
            # result.scale while the next file's "scale" may already have overwritten it (nerf_loader.cu:400-402, :706) -- differing scales race there.
            js = {"n_frames": n_keep, "from_na": False, "scale": 0.4, "frames": frames}
            paths.append(os.path.join(d, f"transforms_{part}.json"))
            json.dump(js, open(paths[-1], "w"))
    else:
        raise KeyError(name)
    return paths


def main(out_path):
    from oracle import ref_loadnerf_host as RL
    out = {"cases": np.array(CASES)}
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None) as root:
        for name in CASES:
            r = RL.load_nerf(build_case(name, root))
            out[f"{name}_focal"] = r["focal"]; out[f"{name}_principal"] = r["principal"]; out[f"{name}_xforms"] = r["xforms"]
            out[f"{name}_pixels"] = np.stack(r["pixels"])
            out[f"{name}_meta"] = np.array([r["scale"], *r["offset"], r["aabb_scale"], int(r["from_na"]), int(r["from_mitsuba"])], np.float64)
            print(name, "images", len(r["pixels"]), "scale", r["scale"], "offset", r["offset"], "aabb_scale", r["aabb_scale"], "from_na", r["from_na"])
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "loader_reference.npz"))
