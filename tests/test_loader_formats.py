"""CPU tests of the file formats either side of the path (SURVEY §8 f): transforms.json loader, config json, mesh writers."""
import json
import os

import numpy as np
import pytest

from neus2_b200 import loader, mesh, synthetic

BASE_JSON = """{
  "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 20000,
     "nested": {"otype": "Adam", "learning_rate": 1e-3, // trailing comment as in configs/nerf/base.json:16
                "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
  "encoding": {"otype": "HashGrid", "n_levels": 14, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16,
               "top_resolution": 2048, "valid_level_scale": 0.02, "base_valid_level_scale": 0.2, "base_training_step": 100},
  "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1},
  "rgb_network": {"otype": "FullyFusedMLP", "activation": "ReLU", "n_neurons": 64, "n_hidden_layers": 2},
  /* block comment */
  "hyperparams": {"first_frame_max_training_step": 2000, "mask_loss_weight": 0.0, "ek_loss_weight": 0.01, "url": "http://x//y"}
}"""


def test_config_json_with_comments():
    js = json.loads(loader._strip_json_comments(BASE_JSON))
    assert js["hyperparams"]["url"] == "http://x//y"          # // inside a string is not a comment
    kw = loader.config_from_json(js)
    assert kw == {"n_levels": 14, "log2_hashmap_size": 19, "base_resolution": 16, "valid_level_scale": 0.02, "base_valid_level_scale": 0.2,
                  "base_training_step": 100, "top_resolution": 2048.0, "ema_decay": 0.95, "learning_rate": 1e-3, "beta1": 0.9, "beta2": 0.99,
                  "epsilon": 1e-15, "l2_reg": 1e-6, "mask_loss_weight": 0.0, "ek_loss_weight": 0.01,
                  # ExponentialDecay around the main Adam; what the json leaves out takes exponential_decay.h's defaults
                  "decay_start": 20000, "decay_interval": 10000, "decay_end": 10000000, "decay_base": 0.1,
                  # no "globalmove" section: the DeltaNetwork's trainer is built from the main optimizer config (testbed.cu:2320-2323)
                  "delta_decay_start": 20000, "delta_decay_interval": 10000, "delta_decay_base": 0.1,
                  "delta_learning_rate": 1e-3, "delta_beta1": 0.9, "delta_beta2": 0.99, "delta_epsilon": 1e-15,
                  # not in this json: the reference's CODE default is true (src/testbed.cu:2123); configs/nerf/base.json:130 sets false
                  "reset_density_grid_after_global_movement": 1}
    off = json.loads(loader._strip_json_comments(BASE_JSON))
    off["hyperparams"]["reset_density_grid_after_global_movement"] = False
    assert loader.config_from_json(off)["reset_density_grid_after_global_movement"] == 0
    gm = json.loads(loader._strip_json_comments(BASE_JSON))
    gm["globalmove"] = {"optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 50, "decay_interval": 25,
                        "decay_base": 0.99, "nested": {"otype": "Adam", "learning_rate": 1e-4, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-10}}}}
    kw = loader.config_from_json(gm)
    assert (kw["delta_learning_rate"], kw["delta_epsilon"], kw["delta_decay_start"], kw["delta_decay_interval"], kw["delta_decay_base"]) == (1e-4, 1e-10, 50, 25, 0.99)
    assert kw["learning_rate"] == 1e-3
    bad = json.loads(loader._strip_json_comments(BASE_JSON))
    bad["network"]["n_neurons"] = 128
    with pytest.raises(RuntimeError):
        loader.config_from_json(bad)


def test_nerf_matrix_conventions():
    rs = np.random.RandomState(0)
    m = rs.normal(size=(3, 4)).astype(np.float32)
    off = np.array([0.5, 0.4, 0.6], dtype=np.float32)
    for kw in (dict(from_na=True), dict(from_na=False), dict(from_mitsuba=True)):
        g = loader.nerf_matrix_to_ngp(m, 0.33, off, **kw)
        back = loader.ngp_matrix_to_nerf(g, 0.33, off, **kw)
        assert np.allclose(back, m, atol=2e-6)
    # from_na: the y/z flips cancel, only the translation is scaled and offset (nerf_loader.h:112-120)
    g = loader.nerf_matrix_to_ngp(m, 0.5, off, from_na=True)
    assert np.array_equal(g[:, :3], m[:, :3]) and np.array_equal(g[:, 3], m[:, 3] * np.float32(0.5) + off)
    # original NeRF convention: flip y and z, then cycle the rows xyz <- yzx
    g = loader.nerf_matrix_to_ngp(m, 1.0, np.zeros(3, np.float32))
    flipped = m * np.array([1, -1, -1, 1], dtype=np.float32)
    assert np.array_equal(g, flipped[[1, 2, 0]])


def test_transforms_json_round_trip(tmp_path):
    data = synthetic.render_sphere_views(n_views=3, width=40, height=24, seed=4, principal=(0.52, 0.47))
    path = loader.write_dataset(str(tmp_path / "scene"), data, scale=0.5, offset=(0.5, 0.5, 0.5), from_na=True)
    got = loader.load_nerf(path, pin=False)
    assert got["from_na"] and got["aabb_scale"] == 1 and got["resolution"] == (40, 24) and got["scale"] == 0.5
    assert np.array_equal(got["pixels"], data["pixels"])
    assert np.array_equal(got["focal"], data["focal"])
    assert np.allclose(got["principal"], data["principal"], atol=1e-7)
    assert np.allclose(got["xforms"], data["xforms"], atol=1e-6)
    # a directory of jsons loads the same thing; n_frames culls; a missing image raises like the reference
    got2 = loader.load_nerf(os.path.dirname(path), pin=False)
    assert np.array_equal(got2["pixels"], got["pixels"])
    js = json.load(open(path))
    js["n_frames"] = 2
    json.dump(js, open(path, "w"))
    assert loader.load_nerf(path, pin=False)["pixels"].shape[0] == 2
    js["frames"][0]["file_path"] = "images/nope.png"
    json.dump(js, open(path, "w"))
    with pytest.raises(RuntimeError):
        loader.load_nerf(path, pin=False)
    with pytest.raises(RuntimeError):
        loader.load_nerf(str(tmp_path / "does_not_exist.json"))


def test_alpha_and_mask_side_images_and_fov_keys(tmp_path):
    from PIL import Image
    d = tmp_path / "s"
    os.makedirs(d)
    rgb = np.full((6, 8, 4), 200, dtype=np.uint8)
    Image.fromarray(rgb, "RGBA").save(d / "a.png")
    alpha = np.zeros((6, 8, 4), dtype=np.uint8)
    alpha[..., 0] = 128
    alpha[..., 3] = 255
    Image.fromarray(alpha, "RGBA").save(d / "a.alpha.png")
    mask = np.zeros((6, 8, 4), dtype=np.uint8)
    mask[0, 0, 0] = 255
    mask[..., 3] = 255
    Image.fromarray(mask, "RGBA").save(d / "dynamic_mask_a.png")
    js = {"camera_angle_x": 0.9, "w": 8, "h": 6, "cx": 4.4, "frames": [{"file_path": "a", "transform_matrix": np.eye(4).tolist()}]}
    json.dump(js, open(d / "transforms.json", "w"))
    got = loader.load_nerf(str(d / "transforms.json"), pin=False)
    px = got["pixels"][0]
    assert tuple(px[0, 0]) == (255, 0, 255, 0)                                    # hot-pink mask pixel
    assert px[1, 1, 3] == np.uint8(255.0 * (((128 / 255.0 + 0.055) / 1.055) ** 2.4))  # alpha side image goes through sRGB -> linear
    f = 0.5 * 8 / np.tan(0.5 * 0.9)
    assert np.allclose(got["focal"][0], [f, f], rtol=1e-6)
    assert np.array_equal(got["principal"][0], [0.5, 0.5])      # "cx" is parsed but never applied by the reference (nerf_loader.cu:431-437, :703 commented out)
    assert not got["from_na"] and abs(got["scale"] - 0.33) < 1e-7 and np.array_equal(got["offset"], [0.5, 0.5, 0.5])


def test_mesh_writers_round_trip(tmp_path):
    V = np.array([[0.5, 0.5, 0.5], [0.75, 0.5, 0.5], [0.5, 0.75, 0.5], [0.5, 0.5, 0.75]], dtype=np.float32)
    N = np.array([[0, 0, 2], [3, 0, 0], [0, 0.5, 0], [0, 0, 0]], dtype=np.float32)
    C = np.array([[0.1, 0.2, 0.3], [1.5, -1, 0.5], [0, 0, 0], [1, 1, 1]], dtype=np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int64)
    p = str(tmp_path / "m.obj")
    mesh.save(p, V, N, C, F, nerf_scale=0.5, nerf_offset=(0.5, 0.5, 0.5))
    v, n, c, f = mesh.load_obj(p)
    assert np.allclose(v, (V - 0.5) / 0.5, atol=1e-5)                 # dataset coordinates
    assert np.allclose(n[:3], [[0, 0, 1], [1, 0, 0], [0, 1, 0]]) and np.allclose(c[1], [1, 0, 0.5])
    assert np.array_equal(f, F[:, ::-1])                              # reversed winding (marching_cubes.cu:953-955)
    p2 = str(tmp_path / "m.ply")
    mesh.save(p2, V, N, C, F)
    lines = open(p2).read().splitlines()
    assert lines[0] == "ply" and "element vertex 4" in lines and "element face 2" in lines and lines[-1] == "3 3 2 0"
