"""Generates tests/golden/render_reference.npz by running the REFERENCE's own tracer kernels (src/testbed_nerf.cu:
init_rays_with_payload_kernel_nerf :2208-2329, advance_pos_nerf :797-846, compact_kernel_nerf :2183-2206,
generate_next_nerf_network_inputs :877-934, composite_kernel_nerf :936-1168) compiled as host code from where they lie
(oracle/Makefile.ref -> oracle/_ref/libref_nerf_host.so), with the host loop of NerfTracer::trace restated around them and the
oracle's network forward (itself pinned against the reference's tiny-cuda-nn kernels) as the network.  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_render_golden.py tests/golden/render_reference.npz
"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
W, H = 40, 30
LATTICES = [((8, 12, 6), (0, 0, 0, 1, 1, 1), (0, 0, 0, 1, 1, 1)), ((12, 8, 8), (0.1, 0.2, 0.0, 0.9, 0.8, 1.0), (0, 0, 0, 1, 1, 1)),
            ((8, 8, 8), (0.25, 0.25, 0.25, 0.75, 0.75, 0.75), (-0.5, -0.5, -0.5, 1.5, 1.5, 1.5))]


def setup():
    from tests import util as U
    spec = importlib.util.spec_from_file_location("make_nerf_kernels_golden", os.path.join(ROOT, "tests", "golden", "make_nerf_kernels_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    data = mk.scene()
    cfg = U.base_cfg()
    params = U.to_half_bits(U.random_params(cfg, seed=2))
    views = []
    for k, (view, spp, snap, min_t) in enumerate([(1, 3, False, 0.01), (2, 0, True, 0.2)]):
        views.append(dict(camera=data["xforms"][view].astype(np.float32), focal=(data["focal"][view] * W / 64).astype(np.float32),
                          screen_center=np.array([0.5, 0.45 + 0.05 * k], np.float32), spp=spp, snap=snap, min_transmittance=min_t))
    return dict(data=data, bitfield=mk.shell_bitfield(data), cfg=cfg, params=params, views=views)


def main(out_path):
    from oracle import oracle as O
    from oracle import ref_nerf_host as RK
    s = setup()

    def net(n, coords, out):
        O.lib().orc_net_forward(C.byref(s["cfg"]), C.c_uint32(n), coords, s["params"].ctypes.data_as(C.c_void_p), out)

    out = {"n_views": np.array(len(s["views"]))}
    for k, v in enumerate(s["views"]):
        r = RK.render_trace(net, W, H, v["spp"], v["focal"], v["camera"], v["screen_center"], s["bitfield"], snap=v["snap"], min_transmittance=v["min_transmittance"])
        for name in ("payload", "rgba", "depth", "n_steps"):
            out[f"v{k}_{name}"] = r[name]
        out[f"v{k}_n_hit"] = np.array(r["n_hit"])
        print("view", k, "alive", int(r["payload"][:, 7].sum()), "hits", r["n_hit"])
    # generate_grid_samples_nerf_uniform (:596-609): the lattice the SDF is evaluated on for marching cubes
    for k, (res, ra, ta) in enumerate(LATTICES):
        pos = np.zeros((res[2], res[1], res[0], 3), np.float32)
        RK.lib().refk_lattice_positions(res[0], res[1], res[2], np.array(ra, np.float32).ctypes.data_as(C.c_void_p), np.array(ta, np.float32).ctypes.data_as(C.c_void_p),
                                        pos.ctypes.data_as(C.c_void_p))
        out[f"lattice{k}"] = pos
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "render_reference.npz"))
