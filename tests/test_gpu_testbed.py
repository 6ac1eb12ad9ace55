"""GPU test of the reference-facing Python surface (neus2_b200.testbed.Testbed mirrors pyngp.Testbed for the NeuS path):
transforms.json -> train -> render -> mesh -> snapshot round trip, with the reference's names and error behaviour."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained(tmp_path_factory):
    from neus2_b200 import loader, synthetic
    from neus2_b200.testbed import Testbed, TestbedMode
    root = tmp_path_factory.mktemp("scene")
    data = synthetic.render_sphere_views(n_views=8, width=96, height=96, seed=0, elevation_deg=(-50.0, 60.0))
    path = loader.write_dataset(str(root), data, scale=0.5, offset=(0.5, 0.5, 0.5), from_na=True)
    tb = Testbed(TestbedMode.Nerf)
    tb.load_training_data(path)
    tb.reload_network_from_json({"encoding": {"otype": "HashGrid", "n_levels": 14}, "hyperparams": {"first_frame_max_training_step": 300}})
    while tb.training_step < 300:
        assert tb.frame()
    yield dict(tb=tb, data=data, root=str(root))


def test_training_reduces_loss_and_stops_at_max_step(trained):
    tb = trained["tb"]
    assert tb.training_step == 300 and 0 < tb.loss < 2e-3
    tb.frame()                                   # frame() past first_frame_max_training_step does not train
    assert tb.training_step == 300


def test_render_training_view_resembles_ground_truth(trained):
    tb, data = trained["tb"], trained["data"]
    tb.snap_to_pixel_centers = True
    tb.background_color = [0.0, 0.0, 0.0, 1.0]
    tb.set_camera_to_training_view(2)
    img = tb.render(96, 96, 1, False)
    assert img.shape == (96, 96, 4) and img.dtype == np.float32
    gt = data["pixels"][2].astype(np.float32) / 255.0
    mse = np.mean((np.clip(img[..., :3], 0, 1) - gt[..., :3] * gt[..., 3:4]) ** 2)
    assert -10 * np.log10(mse) > 25.0
    lin = tb.render(96, 96, 1, True)             # linear=True skips the sRGB transfer
    assert lin[..., :3].mean() < img[..., :3].mean()
    # a free camera in dataset coordinates goes through nerf_matrix_to_ngp like python_api.cu's set_nerf_camera_matrix
    from neus2_b200 import loader
    tb.set_nerf_camera_matrix(loader.ngp_matrix_to_nerf(data["xforms"][2], 0.5, (0.5, 0.5, 0.5), True))
    assert np.allclose(tb.camera_matrix, data["xforms"][2], atol=1e-6)


def test_mesh_and_files(trained):
    tb, root = trained["tb"], trained["root"]
    m = tb.compute_marching_cubes_mesh([100, 100, 100])       # rounded up to 112 like the reference
    V, N, C, F = m["V"], m["N"], m["C"], m["F"]
    assert len(V) > 1000 and V.shape == N.shape == C.shape and F.max() == len(V) - 1
    r = np.linalg.norm(V - 0.5, axis=1)
    assert 0.15 < r.min() and r.max() < 0.42                  # the two-sphere object, roughly
    # vertex normals: area-weighted face normals; the SDF grows outwards and the reference's 'inside' is sdf > 0, so the
    # (un-reversed) winding points inwards and save_mesh flips it
    nn = N / np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-20)
    radial = (V - 0.5) / r[:, None]
    assert np.mean((nn * radial).sum(1) < -0.5) > 0.8
    ref = np.zeros_like(N)
    fa = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 0]] - V[F[:, 2]])
    for k in range(3):
        np.add.at(ref, F[:, k], fa)
    assert np.allclose(N, ref, atol=1e-6)
    assert C.min() >= 0 and C.max() <= 1 and C.std() > 0.01
    out = os.path.join(root, "mesh.obj")
    n_tri = tb.compute_and_save_marching_cubes_mesh(out, [100, 100, 100])
    from neus2_b200 import mesh
    v, n, c, f = mesh.load_obj(out)
    assert len(f) == n_tri == len(F) and np.allclose(v, (V - 0.5) / 0.5, atol=2e-5)


def test_snapshot_round_trip(trained):
    import torch
    from neus2_b200.testbed import Testbed, TestbedMode
    tb, root = trained["tb"], trained["root"]
    path = os.path.join(root, "snap.msgpack")
    tb.save_snapshot(path, True)
    tb2 = Testbed(TestbedMode.Nerf)
    tb2.load_training_data(os.path.join(root, "transforms.json"))
    tb2.load_snapshot(path)
    assert tb2.training_step == tb.training_step
    a, b = tb._ctx, tb2._ctx
    # inference weights survive bit for bit; training weights become the inference ones (Trainer::deserialize, trainer.h:295-304)
    assert torch.equal(a.buffers()[2], b.buffers()[2])
    assert torch.equal(b.buffers()[0].half(), b.buffers()[1])       # fp32 master is the widened fp16 blob
    assert torch.equal(a.optimizer_buffers()[0], b.optimizer_buffers()[0]) and torch.equal(a.optimizer_buffers()[2], b.optimizer_buffers()[2])
    assert a.optimizer_buffers()[3] == b.optimizer_buffers()[3]
    la = a.sdf_lattice((32, 32, 32))
    lb = b.sdf_lattice((32, 32, 32))
    assert torch.equal(la, lb)
    # the occupancy grid travels as fp16: the same cells stay occupied up to threshold ties
    ba = a.debug_buffer("bitfield", "|u1", (128 ** 3 // 8,)).cpu().numpy()
    bb = b.debug_buffer("bitfield", "|u1", (128 ** 3 // 8,)).cpu().numpy()
    assert (np.unpackbits(ba) != np.unpackbits(bb)).mean() < 1e-3
    tb2.first_frame_max_training_step = tb2.training_step + 20
    for _ in range(20):
        tb2.frame()
    assert tb2.training_step == tb.training_step + 20 and tb2.loss < 5e-3
    with pytest.raises(RuntimeError):
        tb2.load_snapshot(os.path.join(root, "missing.msgpack"))
    with pytest.raises(RuntimeError):
        tb2.load_snapshot(_not_a_snapshot(root))


def _not_a_snapshot(root):
    import msgpack
    p = os.path.join(root, "cfg_only.msgpack")
    with open(p, "wb") as f:
        f.write(msgpack.packb({"encoding": {"n_levels": 14}}, use_bin_type=True))
    return p


def test_errors_mirror_the_reference(tmp_path):
    from neus2_b200.testbed import Testbed, TestbedMode
    tb = Testbed(TestbedMode.Nerf)
    with pytest.raises(RuntimeError):
        tb.load_training_data(str(tmp_path / "nope"))
    with pytest.raises(RuntimeError):
        tb.render(8, 8, 1, True)
    with pytest.raises(RuntimeError):
        Testbed(TestbedMode.Sdf)
    with pytest.raises(RuntimeError):
        tb.reload_network_from_file(str(tmp_path / "nope.json"))


def test_sequence_incremental_training(tmp_path):
    """A directory of jsons is a sequence (one json per time frame).  With "predict_global_movement": false the next frame
    continues from the inference weights with a fresh optimizer and trains from its first step
    (m_predict_global_movement_training_step = 0, testbed.cu:2131-2133)."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import run_per_frame
    from neus2_b200.testbed import Testbed, TestbedMode
    run_per_frame.write_synthetic_sequence(str(tmp_path), 2, views=8, res=96)
    tb = Testbed(TestbedMode.Nerf)
    tb.load_training_data(str(tmp_path))
    assert tb.all_training_time_frame == 2 and tb.current_training_time_frame == 0
    cfg = {"optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "Adam", "learning_rate": 1e-3, "after_learning_rate": 5e-4}},
           "hyperparams": {"first_frame_max_training_step": 400, "next_frame_max_training_step": 300, "predict_global_movement": False,
                           "predict_global_movement_training_step": 50}}
    tb.reload_network_from_json(cfg)
    assert tb.predict_global_movement_training_step == 0
    while not (tb.current_training_time_frame == 0 and tb.training_step == 400):
        tb.frame()
    ema0 = tb._ctx.buffers()[2].clone()
    st = tb._ctx.get_state()
    tb.training_network_next_frame()
    assert tb.current_training_time_frame == 1 and tb.training_step == 0
    assert tb._ctx.time_frame() == (1, pytest.approx(5e-4))
    assert torch.equal(tb._ctx.buffers()[1], ema0)       # training weights = previous inference weights
    assert not tb._ctx.global_movement()["use_delta"]
    tb.frame()
    assert tb.training_step == 1 and not torch.equal(tb._ctx.buffers()[1], ema0)   # no frozen phase without global-movement prediction
    while tb.training_step < 300:
        tb.frame()
    tb.frame()
    assert tb.training_step == 300 and tb.current_training_time_frame == 1     # last frame: stops at its budget
    # the mesh follows the object: frame 1's big sphere sits at x = 0.51
    m = tb.compute_marching_cubes_mesh([96, 96, 96])
    big = np.linalg.norm(m["V"] - np.array([0.51, 0.5, 0.5]), axis=1)
    assert np.median(np.abs(big[big < 0.3] - 0.25)) < 4e-3
    assert tb.training_network_next_frame() is False     # "alread last frame"
