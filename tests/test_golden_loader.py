"""CPU pin of neus2_b200.loader.load_nerf against golden vectors produced by the REFERENCE's own dataset loader (src/nerf_loader.cu:
load_nerf, NerfDataset::set_training_image / convert_rgba32, with the reference's json, filesystem and stb_image libraries) compiled as host code
from where it lies (oracle/Makefile.ref, tests/golden/make_loader_golden.py): five small datasets in the dialects of the transforms.json format --
per-frame intrinsic_matrix + from_na (NeuS), camera_angle_x and extension-less paths (NeRF synthetic), fl_x / cx / aabb_scale / scalar offset /
Windows separators (COLMAP exports), .alpha images / dynamic masks / white_ and black_transparent, n_frames culling over several json files.
Intrinsics, camera matrices and the decoded RGBA8 pixels are compared bit for bit."""
import importlib.util
import os

import numpy as np
import pytest

from neus2_b200 import loader
from oracle import ref_loadnerf_host as RL

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "loader_reference.npz")


def _mod():
    spec = importlib.util.spec_from_file_location("make_loader_golden", os.path.join(HERE, "golden", "make_loader_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _mine(paths):
    """Testbed::load_training_data hands load_nerf one json or every json of a directory: here the directory of the case."""
    return loader.load_nerf(paths[0] if len(paths) == 1 else os.path.dirname(paths[0]), pin=False)


@pytest.mark.parametrize("name", _mod().CASES)
def test_loader_matches_the_reference_loader(name, tmp_path):
    g, mod = np.load(GOLD), _mod()
    mine = _mine(mod.build_case(name, str(tmp_path)))
    scale, ox, oy, oz, aabb_scale, from_na, from_mitsuba = g[f"{name}_meta"]
    assert np.float32(mine["scale"]) == np.float32(scale) and np.array_equal(mine["offset"], np.array([ox, oy, oz], np.float32))
    assert mine["aabb_scale"] == int(aabb_scale) and mine["from_na"] == bool(from_na) and not from_mitsuba
    assert np.array_equal(np.asarray(mine["pixels"]), g[f"{name}_pixels"])                          # decode + alpha / mask / transparency handling
    assert np.array_equal(mine["focal"].view(np.uint32), g[f"{name}_focal"].view(np.uint32))
    assert np.array_equal(mine["principal"].view(np.uint32), g[f"{name}_principal"].view(np.uint32))
    assert np.array_equal(mine["xforms"].view(np.uint32), g[f"{name}_xforms"].view(np.uint32))       # nerf_matrix_to_ngp with the file's scale / offset / from_na


def test_fixture_covers_the_special_pixels():
    g = np.load(GOLD)
    px = g["alpha_mask_transparency_pixels"]
    assert (px[:, 0, :3, 3] == 0).all() and (px[:, 1, :3, 3] == 0).all()              # pure white / pure black made transparent
    assert (px[1, 2:5, 4:9] == np.array([255, 0, 255, 0], np.uint8)).all()            # masked pixels: the hot-pink marker with zero alpha
    assert len(np.unique(px[0, 2:, :, 3])) > 20                                       # alpha taken from the .alpha image (sRGB-decoded)
    assert len(g["n_frames_two_jsons_pixels"]) == 3                                   # n_frames 2 + 1 of 3 + 3


@pytest.mark.skipif(not RL.available(), reason="needs the reference tree (oracle/_ref/libref_loadnerf_host.so)")
def test_committed_fixture_is_what_the_reference_source_produces(tmp_path):
    g, mod = np.load(GOLD), _mod()
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name
