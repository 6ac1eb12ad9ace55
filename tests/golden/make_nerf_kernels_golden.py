"""Generates tests/golden/nerf_kernels_reference.npz by running the REFERENCE's own training kernels
(src/testbed_nerf.cu: generate_training_samples_nerf_with_global_movement :1263-1456 and
compute_loss_kernel_train_nerf_with_global_movement :1475-1997, with every device helper they call) compiled as host code from where they
lie into oracle/_ref/libref_nerf_host.so by oracle/Makefile.ref, one simulated thread at a time in ascending order.  CPU only; needs
/root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_nerf_kernels_golden.py tests/golden/nerf_kernels_reference.npz

Two cases: identity global movement, and a non-identity accumulated rotation / transition.  Inputs that are cheap to rebuild from seeds
(the analytic scene's images, the synthetic network outputs) are rebuilt by the test with the functions below; the occupancy bitfield is stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

N_RAYS, MAX_SAMPLES, N_RAYS_TOTAL, B = 640, 1 << 16, 777 * 4096 + 5, 4096


LOSS_VARIANTS = [dict(loss_scale=128.0, bg=(0.3, 0.6, 0.9), random_bg=False, mask_w=0.0, ek_w=0.01, cos_anneal=1.0),
                 dict(loss_scale=64.0, bg=(0, 0, 0), random_bg=True, mask_w=0.05, ek_w=0.1, cos_anneal=0.35)]


def scene():
    from neus2_b200 import synthetic
    return synthetic.render_sphere_views(n_views=4, width=64, height=48, seed=3)


def shell_bitfield(data):
    """occupancy of the analytic surface shell, mip 0 only, packed like the reference's bitfield (morton order, 8 cells per byte)"""
    from neus2_b200 import synthetic
    G = 128
    idx = np.arange(G ** 3, dtype=np.uint32)

    def dec(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    pts = (np.stack([dec(idx), dec(idx >> 1), dec(idx >> 2)], axis=1).astype(np.float64) + 0.5) / G
    occ = np.abs(synthetic.sphere_sdf(pts, data["spheres"])) < 3.0 / G
    bits = np.packbits(occ.astype(np.uint8), bitorder="little")
    out = np.zeros(G ** 3 // 8 * 8, dtype=np.uint8)
    out[: len(bits)] = bits
    return out


def network_outputs(rs, n_kept, numsteps, coords, max_samples):
    """plausible 16-wide outputs: raw colours, an sdf that crosses zero along the ray, normals of varying length, variance, view dir"""
    out = np.zeros((max_samples, 16), np.float16)
    for r in range(n_kept):
        c, b = int(numsteps[r, 0]), int(numsteps[r, 1])
        frac = np.linspace(0, 1, c, dtype=np.float32)
        out[b:b + c, 0:3] = rs.normal(0, 1.5, (c, 3))
        out[b:b + c, 3] = (0.25 - 0.6 * frac + rs.normal(0, 0.01, c)) * (1 if r % 3 else 0.2)
        nrm = rs.normal(0, 1, (c, 3))
        nrm = nrm / np.linalg.norm(nrm, axis=1, keepdims=True) * rs.uniform(0.7, 1.3, (c, 1))
        out[b:b + c, 4:7] = nrm
        out[b:b + c, 7] = 0.3 + 0.02 * (r % 5)
        out[b:b + c, 8:11] = coords[b:b + c, 4:7]
    return np.ascontiguousarray(out.view(np.uint16))


def _rows_of(compacted, coords):
    """coords_compacted is a gather of coords rows (the reference copies whole NerfCoordinates): store the row indices instead of the floats"""
    key = {c.tobytes(): i for i, c in enumerate(coords)}
    return np.array([key[c.tobytes()] for c in compacted], dtype=np.uint32)


def main(out_path):
    from oracle import ref_nerf_host as RK
    data = scene()
    bitfield = shell_bitfield(data)
    out = {"bitfield_packed": np.packbits(np.unpackbits(bitfield[: 128 ** 3 // 8]))}    # mip 0 only, the rest is zero
    a = np.deg2rad(3.0)
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    cases = [dict(rot=RK.IDENTITY_ROT_H, tr=RK.ZERO_TRANS_H, seed=1337, mask_w=0.0),
             dict(rot=np.concatenate([Rz.reshape(9), [0, 0, 0]]).astype(np.float16).view(np.uint16),
                  tr=np.array([0.02, -0.01, 0.015, 0], np.float16).view(np.uint16), seed=99, mask_w=0.1)]
    for k, cs in enumerate(cases):
        rng_state = np.array([0, 0], np.uint64)
        from oracle import oracle as O
        r = O.Pcg32(seed=cs["seed"]); r.advance()      # pcg32 itself is pinned by tests/test_golden_ld.py
        rng_state[:] = r.st
        g = RK.generate_training_samples(N_RAYS, MAX_SAMPLES, N_RAYS_TOTAL, rng_state, data, bitfield, rotation=cs["rot"], transition=cs["tr"])
        kk = g["n_kept"]
        n = int(g["numsteps"][kk - 1].sum())
        rs = np.random.RandomState(10 + k)
        net = network_outputs(rs, kk, g["numsteps"], g["coords"], MAX_SAMPLES)
        lo = RK.compute_loss(N_RAYS, kk, N_RAYS_TOTAL, rng_state, B, data, net, g["ray_indices"], g["rays"], g["numsteps"], g["coords"],
                             mask_loss_weight=cs["mask_w"], rotation=cs["rot"], transition=cs["tr"])
        m = min(lo["n_compacted"], B)
        pre = f"c{k}_"
        out.update({pre + "rng_state": rng_state, pre + "rot": cs["rot"], pre + "tr": cs["tr"], pre + "mask_w": np.float32(cs["mask_w"]),
                    pre + "n_kept": np.array(kk), pre + "n_samples": np.array(g["n_samples"]), pre + "ray_indices": g["ray_indices"][:kk],
                    pre + "rays": g["rays"][:kk], pre + "numsteps": g["numsteps"][:kk], pre + "coords": g["coords"][:n],
                    pre + "n_compacted": np.array(lo["n_compacted"]), pre + "numsteps_compacted": lo["numsteps"][:kk], pre + "dloss": lo["dloss"][:m],
                    pre + "coords_compacted_rows": _rows_of(lo["coords_compacted"][:m], g["coords"][:n]), pre + "loss": lo["loss"][:kk], pre + "ek_loss": lo["ek_loss"][:kk],
                    pre + "mask_loss": lo["mask_loss"][:kk]})
        print("case", k, "kept", kk, "samples", g["n_samples"], "compacted", lo["n_compacted"])
    # loss-stage variants on case 0's samples: fixed background colour, cosine annealing in progress, another loss scale / eikonal weight
    g0 = {k[3:]: v for k, v in out.items() if k.startswith("c0_")}
    coords = np.zeros((MAX_SAMPLES, 7), np.float32); coords[: len(g0["coords"])] = g0["coords"]
    kk = int(g0["n_kept"])
    net = network_outputs(np.random.RandomState(10), kk, g0["numsteps"], coords, MAX_SAMPLES)
    ri = np.zeros(N_RAYS, np.uint32); ri[:kk] = g0["ray_indices"]
    rays = np.zeros((N_RAYS, 6), np.float32); rays[:kk] = g0["rays"]
    ns = np.zeros((N_RAYS, 2), np.uint32); ns[:kk] = g0["numsteps"]
    for k, v in enumerate(LOSS_VARIANTS):
        lo = RK.compute_loss(N_RAYS, kk, N_RAYS_TOTAL, g0["rng_state"], B, data, net, ri, rays, ns, coords, loss_scale=v["loss_scale"], background_color=v["bg"],
                             random_bg_color=v["random_bg"], mask_loss_weight=v["mask_w"], ek_loss_weight=v["ek_w"], cos_anneal_ratio=v["cos_anneal"])
        m = min(lo["n_compacted"], B)
        pre = f"l{k}_"
        out.update({pre + "n_compacted": np.array(lo["n_compacted"]), pre + "numsteps_compacted": lo["numsteps"][:kk], pre + "dloss": lo["dloss"][:m],
                    pre + "loss": lo["loss"][:kk], pre + "ek_loss": lo["ek_loss"][:kk], pre + "mask_loss": lo["mask_loss"][:kk]})
        print("loss variant", k, "compacted", lo["n_compacted"])
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "nerf_kernels_reference.npz"))
