"""ctypes binding of libneus2_b200.so (the C ABI in include/neus2_b200.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing or a call fails this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NEUS2_B200_LIB") or os.path.join(_HERE, "libneus2_b200.so")   # the override selects an experimental build of the SAME library


class Neus2Error(RuntimeError):
    """Mirrors the std::runtime_error -> RuntimeError mapping of the reference's pybind11 module."""


class Config(C.Structure):
    _fields_ = [
        ("n_levels", C.c_uint32), ("log2_hashmap_size", C.c_uint32), ("base_resolution", C.c_uint32),
        ("top_resolution", C.c_float), ("per_level_scale", C.c_float), ("valid_level_scale", C.c_float),
        ("base_valid_level_scale", C.c_float), ("base_training_step", C.c_uint32),
        ("sdf_bias", C.c_float),
        ("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
        ("l2_reg", C.c_float), ("ema_decay", C.c_float),
        ("mask_loss_weight", C.c_float), ("ek_loss_weight", C.c_float), ("anneal_end", C.c_int32),
        ("batch_size", C.c_uint32),
        ("aabb_scale", C.c_int32), ("seed", C.c_uint32), ("random_bg_color", C.c_int32), ("snap_to_pixel_centers", C.c_int32),
        ("background_color", C.c_float * 3), ("density_grid_decay", C.c_float), ("loss_scale", C.c_float),
        ("world_size", C.c_uint32), ("max_rays_per_batch", C.c_uint32), ("forward_early_termination", C.c_int32), ("prefetch_next_samples", C.c_int32),
        ("delta_learning_rate", C.c_float), ("delta_beta1", C.c_float), ("delta_beta2", C.c_float), ("delta_epsilon", C.c_float),
        ("delta_decay_start", C.c_uint32), ("delta_decay_interval", C.c_uint32), ("delta_decay_base", C.c_float),
        ("decay_start", C.c_uint32), ("decay_interval", C.c_uint32), ("decay_end", C.c_uint32), ("decay_base", C.c_float),
        ("reset_density_grid_after_global_movement", C.c_int32),
    ]


class StepStats(C.Structure):
    _fields_ = [
        ("training_step", C.c_uint32), ("rays_per_batch", C.c_uint32), ("rays_per_batch_next", C.c_uint32),
        ("n_rays_kept", C.c_uint32), ("measured_batch_size_before_compaction", C.c_uint32), ("measured_batch_size", C.c_uint32),
        ("loss", C.c_float), ("ek_loss", C.c_float), ("mask_loss", C.c_float), ("loss_valid", C.c_int32),
        ("density_grid_updated", C.c_int32), ("first_pass_samples", C.c_uint32), ("samples_prefetched", C.c_int32),
    ]


class RenderParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("focal_length", C.c_float * 2), ("screen_center", C.c_float * 2),
        ("camera_matrix", C.c_float * 12), ("sample_index", C.c_uint32), ("snap_to_pixel_centers", C.c_int32),
        ("min_transmittance", C.c_float), ("background_color", C.c_float * 4), ("exposure", C.c_float), ("to_srgb", C.c_int32),
        ("use_inference_params", C.c_int32), ("tile_x0", C.c_int32), ("tile_y0", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32),
        ("row_block", C.c_int32), ("row_stride", C.c_int32),
    ]


EXPORTS = [
    "neus2_last_error", "neus2_default_config", "neus2_learning_rate_factor", "neus2_next_rays_per_batch", "neus2_ctx_create", "neus2_ctx_destroy", "neus2_n_params", "neus2_param_layout",
    "neus2_grid_tables", "neus2_buffers", "neus2_init_params", "neus2_set_params", "neus2_get_params", "neus2_set_training_step",
    "neus2_forward", "neus2_sdf", "neus2_backward", "neus2_optimizer_step", "neus2_set_dataset", "neus2_update_density_grid",
    "neus2_train_step_begin", "neus2_train_step_end", "neus2_local_counts", "neus2_train_step", "neus2_debug_buffer",
    "neus2_get_state", "neus2_set_state", "neus2_set_bitfield", "neus2_launch_count", "neus2_set_profiling", "neus2_get_stage_times",
    "neus2_default_render_params", "neus2_render", "neus2_render_stats", "neus2_sdf_lattice", "neus2_marching_cubes", "neus2_mesh_buffers", "neus2_counts_device",
    "neus2_ipc_export", "neus2_ipc_attach", "neus2_optimizer_shard", "neus2_optimizer_step_sharded",
    "neus2_optimizer_buffers", "neus2_set_optimizer_step", "neus2_set_density_grid", "neus2_mesh_attributes", "neus2_get_counters", "neus2_next_frame", "neus2_get_time_frame",
    "neus2_train_step_counts_ready", "neus2_reinit_sdf_mlp", "neus2_begin_global_movement", "neus2_set_delta_mode", "neus2_get_global_movement", "neus2_set_global_movement", "neus2_delta_gradients_device",
    "neus2_loss_scalars_device", "neus2_peer_exchange", "neus2_peer_barrier", "neus2_canonical_training_step", "neus2_set_debug_bits", "neus2_get_stage_times_ext",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Neus2Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C neus2_b200/csrc`). neus2_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.neus2_last_error.restype = C.c_char_p
    lib.neus2_n_params.restype = C.c_uint64
    lib.neus2_learning_rate_factor.restype = C.c_float
    lib.neus2_learning_rate_factor.argtypes = [C.c_void_p, C.c_uint32]
    lib.neus2_next_rays_per_batch.restype = C.c_uint32
    lib.neus2_next_rays_per_batch.argtypes = [C.c_uint32] * 5
    lib.neus2_canonical_training_step.restype = C.c_uint32
    lib.neus2_canonical_training_step.argtypes = [C.c_uint32] * 3
    lib.neus2_launch_count.restype = C.c_uint64
    lib.neus2_n_params.argtypes = [C.c_void_p]
    lib.neus2_launch_count.argtypes = [C.c_void_p]
    lib.neus2_ctx_destroy.argtypes = [C.c_void_p]
    lib.neus2_ctx_destroy.restype = None
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise Neus2Error(load().neus2_last_error().decode())
