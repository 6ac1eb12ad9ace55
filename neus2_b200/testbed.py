"""`Testbed` -- the Python surface the reference's drivers use (`pyngp.Testbed(TestbedMode.Nerf)`, src/python_api.cu:216-614),
restated over the C ABI for the NeuS path: same method / property names, argument meaning and error behaviour
(RuntimeError) for the members that scripts/run.py, scripts/run_dynamic.py and the evaluation scripts touch.  Everything
numerical happens in libneus2_b200.so; this file is host logic only (state, camera set-up, file formats).

Not mirrored (out of scope): GUI / window methods, the Sdf / Image / Volume modes, DLSS, camera paths,
per-image latents, envmap / distortion / exposure optimisation, and the dynamic-scene (delta network) members.
"""
import enum
import os

import numpy as np

from . import Context, default_config, loader, snapshot as _snapshot, mesh as _mesh


def _rotation_6d_to_matrix(r6):
    """rotation_6d_to_matrix (common_operation.cuh:37-60): Gram-Schmidt of the two 3-vectors, columns b1, b2, b1 x b2."""
    a1, a2 = np.asarray(r6[:3], np.float32), np.asarray(r6[3:6], np.float32)
    b1 = a1 / np.linalg.norm(a1)
    u2 = a2 - np.dot(b1, a2) * b1
    b2 = u2 / np.linalg.norm(u2)
    return np.stack([b1, b2, np.cross(b1, b2)], axis=1).astype(np.float32)


class TestbedMode(enum.Enum):
    Nerf = 0
    Sdf = 1
    Image = 2
    Volume = 3


class ColorSpace(enum.Enum):
    Linear = 0
    SRGB = 1


class _Training:
    def __init__(self):
        self.near_distance = 0.0          # testbed.h: nerf.training.near_distance (unused by the NeuS sample generator, which starts at the box)
        self.random_bg_color = True
        self.linear_colors = False
        self.n_images_for_training = 0


class _Nerf:
    def __init__(self):
        self.training = _Training()
        self.rendering_min_transmittance = 0.01
        self.cone_angle_constant = 0.0
        self.render_with_camera_distortion = False


class Testbed:
    def __init__(self, mode=TestbedMode.Nerf, device=0):
        if mode not in (TestbedMode.Nerf, "nerf"):
            raise RuntimeError("only TestbedMode.Nerf is built")
        self.device = device
        self.nerf = _Nerf()
        self.shall_train = True
        self.background_color = [0.0, 0.0, 0.0, 1.0]
        self.snap_to_pixel_centers = False          # python default; run.py sets True for screenshots (testbed.h:623)
        self.color_space = ColorSpace.Linear
        self.exposure = 0.0
        self.fov_axis = 1
        self.zoom = 1.0
        self.screen_center = np.array([0.5, 0.5], dtype=np.float32)
        self.relative_focal_length = np.array([1.0, 1.0], dtype=np.float32)
        self.camera_matrix = np.array([[1, 0, 0, 0.5], [0, -1, 0, 0.5], [0, 0, -1, 2.0]], dtype=np.float32)
        self.fov = 50.625
        self.first_frame_max_training_step = 2000
        self.next_frame_max_training_step = 2000
        self.predict_global_movement = False
        self.predict_global_movement_training_step = 50
        self.finetune_global_movement = True
        self.incremental_reinit_sdf_mlp = False
        self.incremental_reinit_sdf_mlp_iters = 10
        self.after_learning_rate = 0.0
        self.current_training_time_frame = 0
        self.all_training_time_frame = 1
        self._json_paths = []
        self.loss = 0.0
        self.ek_loss = 0.0
        self.mask_loss = 0.0
        self._network_config = {}
        self._cfg_kwargs = {}
        self._data = None
        self._ctx = None
        self._mesh = None

    # ---- configuration / data --------------------------------------------------------------------------------------
    def reload_network_from_file(self, path, output_path=""):
        """Testbed::reload_network_from_file (testbed.cu:2351-2366): json with comments, or a msgpack snapshot."""
        if not os.path.exists(path):
            raise RuntimeError(f"Network config path '{path}' does not exist.")
        if path.endswith(".msgpack"):
            self._network_config = _snapshot.read(path)
        else:
            self._network_config = loader.load_json(path)
        self.reload_network_from_json(self._network_config)

    def reload_network_from_json(self, cfg_json):
        self._network_config = cfg_json
        self._cfg_kwargs = loader.config_from_json(cfg_json)
        hp = cfg_json.get("hyperparams", {})
        self.first_frame_max_training_step = int(hp.get("first_frame_max_training_step", self.first_frame_max_training_step))
        self.next_frame_max_training_step = int(hp.get("next_frame_max_training_step", self.next_frame_max_training_step))
        self.predict_global_movement = bool(hp.get("predict_global_movement", False))          # testbed.cu:2129
        self.predict_global_movement_training_step = int(hp.get("predict_global_movement_training_step", 300))   # :2130
        if not self.predict_global_movement:
            self.predict_global_movement_training_step = 0                                      # :2131-2133
        self.finetune_global_movement = bool(hp.get("finetune_global_movement", True))          # :2125
        self.incremental_reinit_sdf_mlp = bool(hp.get("incremental_reinit_sdf_mlp", False))
        self.incremental_reinit_sdf_mlp_iters = max(1, int(hp.get("incremental_reinit_sdf_mlp_iters", 10)))
        leaf = cfg_json.get("optimizer", {})
        while "nested" in leaf:
            leaf = leaf["nested"]
        self.after_learning_rate = float(leaf.get("after_learning_rate", leaf.get("learning_rate", 1e-3)))
        self.nerf.training.linear_colors = bool(hp.get("linear_color", False))
        if self.nerf.training.linear_colors:
            raise RuntimeError("train_in_linear_colors is not built (base.json trains in sRGB)")
        self._reset_network()

    def load_training_data(self, path):
        """Testbed::load_training_data (testbed.cu:196-262) for Nerf mode."""
        # a directory of jsons is a SEQUENCE: one json per time frame, the first one is loaded now (Testbed::load_nerf, testbed_nerf.cu:2964-3002)
        self._json_paths = [p for p in loader.list_json_paths(path) if "downsample" not in os.path.basename(p)]
        self.all_training_time_frame = len(self._json_paths)
        self.current_training_time_frame = 0
        self._data = loader.load_nerf(self._json_paths[0])
        if self._data["aabb_scale"] != 1:
            raise RuntimeError("aabb_scale != 1 is not built (NeuS2 object-centric scenes use 1)")
        self.nerf.training.n_images_for_training = len(self._data["paths"])
        self._reset_network()
        self.set_camera_to_training_view(0)

    def _reset_network(self):
        """Testbed::reset_network (testbed.cu:2084-2349): new parameters, optimizer state and occupancy grid."""
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
        if self._data is None:
            return
        kw = dict(self._cfg_kwargs)
        kw["random_bg_color"] = int(bool(self.nerf.training.random_bg_color))
        kw["background_color"] = self.background_color[:3]
        kw["snap_to_pixel_centers"] = 1                       # training always snaps (testbed.h:623 default for the trainer)
        self._ctx = Context(default_config(**kw), device=self.device)
        self._ctx.init_params()
        d = self._data
        self._ctx.set_dataset(d["pixels"], d["focal"], d["principal"], d["xforms"])
        self.loss = self.ek_loss = self.mask_loss = 0.0

    def _need_ctx(self):
        if self._ctx is None:
            raise RuntimeError("no training data loaded: call load_training_data first")
        return self._ctx

    # ---- training --------------------------------------------------------------------------------------------------
    @property
    def training_step(self):
        return 0 if self._ctx is None else self._ctx.get_state()["training_step"]

    def _max_step(self):
        return self.first_frame_max_training_step if self.current_training_time_frame == 0 else self.next_frame_max_training_step

    def frame(self):
        """Testbed::frame (testbed.cu:1700-1850) without a window: when the current time frame has had its steps, move on to the
        next one (testbed.cu:1750-1756), then one training step while shall_train.  Unlike the reference, which leaves stopping to
        its scripts, the last frame simply stops training at its step budget.  Returns True."""
        if self._ctx is None or not self.shall_train:
            return True
        if self.training_step >= self._max_step():
            if self.current_training_time_frame >= self.all_training_time_frame - 1:
                return True
            self.training_network_next_frame()
        self.train(1 << 18)
        return True

    def training_network_next_frame(self):
        """Testbed::training_network_next_frame (testbed.cu:2001-2082).  With "predict_global_movement" the delta trained on the
        frame that just ended is folded into the accumulated rigid motion and a fresh DeltaNetwork (identity, own Adam) is trained
        alone for predict_global_movement_training_step steps, then together with the canonical model (:2054, 2076-2078, 2659-2667)."""
        if self.current_training_time_frame >= self.all_training_time_frame - 1:
            return False                                              # "alread last frame"
        ctx = self._need_ctx()
        self.current_training_time_frame += 1
        self._data = loader.load_nerf(self._json_paths[self.current_training_time_frame])
        d = self._data
        ctx.set_dataset(d["pixels"], d["focal"], d["principal"], d["xforms"])
        if self.predict_global_movement:
            ctx.begin_global_movement()
        ctx.next_frame(after_learning_rate=self.after_learning_rate, frozen_steps=self.predict_global_movement_training_step)
        if self.incremental_reinit_sdf_mlp and self.current_training_time_frame % self.incremental_reinit_sdf_mlp_iters == 0:
            ctx.reinit_sdf_mlp()                                      # testbed.cu:2065-2068
        return True

    def change_to_frame(self, i):
        """python_api.cu `change_to_frame`: load the images / cameras of time frame i (for rendering and evaluation)."""
        if not 0 <= i < self.all_training_time_frame:
            raise RuntimeError("frame index out of range")
        self._data = loader.load_nerf(self._json_paths[i])

    def train(self, batch_size=1 << 18):
        ctx = self._need_ctx()
        if int(batch_size) != int(ctx.cfg.batch_size):
            raise RuntimeError(f"the context was created for batch_size {ctx.cfg.batch_size}")
        if (self.predict_global_movement and not self.finetune_global_movement and self.current_training_time_frame >= 1
                and self.training_step == self.predict_global_movement_training_step):
            ctx.set_delta_mode(False)                                 # m_train_delta = false (:2659-2663): the delta leaves the network
        st = ctx.train_step()
        if st.measured_batch_size == 0:
            # train_nerf (testbed_nerf.cu:3540-3546): "Nerf training generated 0 samples. Aborting training."
            self.loss = self.ek_loss = self.mask_loss = 0.0
            self.shall_train = False
            return st
        if st.loss_valid:
            self.loss, self.ek_loss, self.mask_loss = float(st.loss), float(st.ek_loss), float(st.mask_loss)
        return st

    # ---- camera ----------------------------------------------------------------------------------------------------
    def set_nerf_camera_matrix(self, cam):
        """python_api.cu: m_camera = dataset.nerf_matrix_to_ngp(cam) -- `cam` is camera-to-world in the dataset's convention."""
        d = self._data or dict(scale=float(loader.NERF_SCALE), offset=np.array([0.5, 0.5, 0.5], np.float32), from_na=False)
        self.camera_matrix = loader.nerf_matrix_to_ngp(np.asarray(cam, dtype=np.float32)[:3, :4], d["scale"], d["offset"], d["from_na"])

    def set_camera_to_training_view(self, trainview):
        """testbed.cu:264-270."""
        d = self._data
        if d is None:
            raise RuntimeError("no training data loaded")
        res = d["resolution"]
        self.camera_matrix = d["xforms"][trainview].copy()
        self.relative_focal_length = d["focal"][trainview] / np.float32(res[self.fov_axis])
        self.screen_center = np.float32(1.0) - d["principal"][trainview]

    @property
    def fov(self):
        """Testbed::fov (testbed.cu:1829-1831): focal_length_to_fov(1, relative_focal_length[fov_axis])."""
        return float(2.0 * 180.0 / np.pi * np.arctan(0.5 / float(self.relative_focal_length[self.fov_axis])))

    @fov.setter
    def fov(self, degrees):
        f = loader.fov_to_focal_length(1, degrees)
        self.relative_focal_length = np.array([f, f], dtype=np.float32)

    def _render_params(self, width, height):
        ctx = self._need_ctx()
        f = self.relative_focal_length * np.float32((width, height)[self.fov_axis]) * np.float32(self.zoom)      # calc_focal_length
        sc = (np.float32((np.float32(0.5) - self.screen_center[0]) * np.float32(self.zoom) + np.float32(0.5)),
              np.float32((0.5 - float(self.screen_center[1])) * float(np.float32(self.zoom)) + 0.5))             # render_screen_center
        return ctx.render_params(width, height, camera_matrix=self.camera_matrix, focal_length=f, screen_center=sc,
                                 snap_to_pixel_centers=int(bool(self.snap_to_pixel_centers)), min_transmittance=float(self.nerf.rendering_min_transmittance),
                                 background_color=self.background_color, exposure=float(self.exposure))

    def render(self, width, height, spp=1, linear=True):
        """Testbed::render_to_cpu (python_api.cu:123-170): float32 [height, width, 4]; sRGB-encoded unless `linear`."""
        rp = self._render_params(int(width), int(height))
        rp.to_srgb = 0 if linear else 1
        return self._ctx.render(rp, spp=int(spp)).cpu().numpy()

    # ---- mesh ------------------------------------------------------------------------------------------------------
    def compute_marching_cubes_mesh(self, resolution=(256, 256, 256), aabb=None, thresh=0.0):
        """Testbed::compute_marching_cubes_mesh (python_api.cu:99-121) -> {"V", "N", "C", "F"} numpy arrays."""
        ctx = self._need_ctx()
        res = [(int(r) + 15) // 16 * 16 for r in np.broadcast_to(np.asarray(resolution), (3,))]      # testbed_nerf.cu:4176-4178
        bb = (0, 0, 0, 1, 1, 1) if aabb is None else tuple(float(v) for v in np.asarray(aabb).reshape(-1))
        self._mesh = _mesh.extract(ctx, res, bb, float(thresh))
        out = {k: v.cpu().numpy() for k, v in self._mesh.items()}
        gm = ctx.global_movement()
        if not (np.array_equal(gm["rotation"], np.eye(3, dtype=np.float32)) and not gm["transition"].any()):
            # transform_mesh_with_6d (testbed_nerf.cu:108-135, 4190-4196): v <- R_acc^-1 (v - t_acc), back into the frame's world space.
            # (The reference shades the moved vertices; here normals / colours are taken at the extracted positions and the normals rotated.)
            Rinv = np.linalg.inv(gm["rotation"].astype(np.float64))
            out["V"] = ((out["V"].astype(np.float64) - gm["transition"].astype(np.float64)) @ Rinv.T).astype(np.float32)
            out["N"] = (out["N"].astype(np.float64) @ Rinv.T).astype(np.float32)
        return out

    def save_transform(self, path):
        """Testbed::save_transform (testbed.cu:3117-3142): the rigid motion a snapshot taken now would hold (delta o accumulated,
        save_global_movement_rotation_6d_kernel, common_operation.cuh:582-617), as 3 rows of the rotation + the transition."""
        R, t = self.global_transform()
        with open(path, "w") as f:
            for r in R:
                f.write("%f %f %f\n" % tuple(float(v) for v in r))
            f.write("%f %f %f\n\n" % tuple(float(v) for v in t))

    def global_transform(self):
        """(R [3,3], t [3]) = delta o accumulated: ray_canonical = R ray + t (positions additionally through the per-sample delta)."""
        gm = self._need_ctx().global_movement()
        Rd = _rotation_6d_to_matrix(gm["delta_rotation_6d"].astype(np.float16).astype(np.float32))
        td = gm["delta_transition"].astype(np.float16).astype(np.float32)
        R = (Rd @ gm["rotation"]).astype(np.float16).astype(np.float32)
        t = (Rd @ (gm["transition"] + td)).astype(np.float16).astype(np.float32)
        return R, t

    def compute_and_save_marching_cubes_mesh(self, filename, resolution=(256, 256, 256), aabb=None, thresh=0.0, generate_uvs_for_obj_file=False):
        """Testbed::compute_and_save_marching_cubes_mesh (testbed.cu:308-316): OBJ or PLY in DATASET coordinates."""
        if generate_uvs_for_obj_file:
            raise RuntimeError("uv unwrapping is not built")
        m = self.compute_marching_cubes_mesh(resolution, aabb, thresh)
        d = self._data
        _mesh.save(filename, m["V"], m["N"], m["C"], m["F"], d["scale"], d["offset"])
        return len(m["F"])

    # ---- snapshots -------------------------------------------------------------------------------------------------
    def save_snapshot(self, path, include_optimizer_state=False):
        _snapshot.save(self._need_ctx(), path, self._network_config, include_optimizer_state, loss=self.loss,
                       aabb_scale=1 if self._data is None else self._data["aabb_scale"])

    def load_snapshot(self, path):
        if not os.path.exists(path):
            raise RuntimeError(f"File '{path}' does not exist.")
        cfg = _snapshot.read(path)
        if "snapshot" not in cfg:
            raise RuntimeError(f"File '{path}' does not contain a snapshot.")
        self._network_config = {k: v for k, v in cfg.items() if k != "snapshot"}
        self._cfg_kwargs = loader.config_from_json(self._network_config)
        self._reset_network()
        self.loss = _snapshot.restore(self._need_ctx(), cfg["snapshot"])
