"""Generates tests/golden/ref_kernels.npz by running the REFERENCE's own kernels (dependencies/my_tcnn, compiled from the
reference sources in place into oracle/_ref/libtcnn_ref.so) on seeded inputs.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_kernels.npz'    # then copy into tests/golden/

The committed .npz lets the CPU-only test-suite pin the oracle (oracle/neus_oracle.cpp) against real reference outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(out_path):
    import torch
    from oracle import oracle as O
    from oracle import ref_kernels as RK
    from tests import util as U

    cfg = U.base_cfg()
    lay = O.param_layout(cfg)
    rs = np.random.RandomState(2024)
    out = {}

    # ---- hash grid (L=14, F=2, T=2^19): only the entries the samples touch are stored (sparse golden) ----
    n = 256
    coords = U.random_coords(n, seed=11)
    coords[:8, :3] = np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [0.999999, 0.0, 1.0], [1e-7, 0.25, 0.75], [0.3333333, 0.6666667, 0.1],
                               [0.0625, 0.0625, 0.0625], [0.9, 0.1, 0.5]], dtype=np.float32)      # cell-boundary / edge positions
    grid = rs.uniform(-0.05, 0.05, lay["n_grid"]).astype(np.float16)
    dL_dy = rs.normal(0, 1e-2, (28, n)).astype(np.float16)
    v = rs.normal(0, 1e-2, (n, 3)).astype(np.float32)
    g = RK.RefGrid(per_level_scale=cfg.per_level_scale_py)
    r = g.eval(torch.from_numpy(coords).cuda(), torch.from_numpy(grid).cuda(), torch.from_numpy(dL_dy).cuda(), torch.from_numpy(v).cuda(),
               torch.from_numpy(dL_dy).cuda())
    grad = r["grad"].cpu().numpy()
    # sparse golden: every grid entry the samples touch (8 corners x 14 levels x 2 features), from the index function
    off, res, sc = O.cfg_tables(cfg)
    touched = set()
    for l in range(cfg.n_levels):
        hs = int(off[l + 1] - off[l])
        pf = (coords[:, :3] * np.float32(sc[l]) + np.float32(0.5)).astype(np.float32)
        cell = np.floor(pf).astype(np.int64)
        for c in range(8):
            for a, b, cz in cell + np.array([c & 1, (c >> 1) & 1, (c >> 2) & 1]):
                i = O.grid_index(hs, int(res[l]), int(a), int(b), int(cz)) + 2 * int(off[l])
                touched.add(i); touched.add(i + 1)
    nz = np.array(sorted(touched), dtype=np.uint32)
    assert not np.setdiff1d(np.nonzero(grad.view(np.uint16))[0], nz).size, "reference wrote a gradient outside the predicted corner set"
    out.update(grid_seed=np.uint32(2024), grid_coords=coords, grid_dL_dy=dL_dy, grid_v=v, grid_touched=nz, grid_values=grid[nz],
               grid_enc=r["enc"].cpu().numpy(), grid_dL_dx=r["dL_dx"].cpu().numpy(), grid_dL_ddLdy=r["dL_ddLdy"].cpu().numpy(),
               grid_grad=grad[nz], per_level_scale=np.float32(cfg.per_level_scale_py))
    # progressive level mask (set_training_step, grid.h:2427-2440)
    g.set_training_step(150)
    r2 = g.eval(torch.from_numpy(coords).cuda(), torch.from_numpy(grid).cuda(), want_grad=False, which=1)
    out["grid_enc_step150"] = r2["enc"].cpu().numpy()

    # ---- fused MLPs: sdf (32 -> 64 -> 16, SoA input) and colour (48 -> 64 -> 64 -> 16) ----
    for name, n_in, hl, soa in [("sdf", 32, 1, True), ("rgb", 48, 2, False)]:
        m = 256
        n_params = 64 * n_in + (hl - 1) * 4096 + 1024
        W = rs.uniform(-0.25, 0.25, n_params).astype(np.float16)
        x = rs.uniform(-1, 1, (m, n_in)).astype(np.float16)
        dL = np.zeros((m, 16), dtype=np.float16)
        dL[:, :4] = rs.normal(0, 1e-2, (m, 4)).astype(np.float16)
        d2 = rs.normal(0, 1e-2, (m, n_in)).astype(np.float16)
        seed = np.zeros((m, 16), dtype=np.float16)
        seed[:, 0] = 1
        net = RK.RefMlp(n_in, hl)
        xin = torch.from_numpy(np.ascontiguousarray(x.T) if soa else x).cuda()
        rr = net.eval(xin, torch.from_numpy(W).cuda(), in_soa=soa, dL_dout=torch.from_numpy(dL).cuda(), d_in2=torch.from_numpy(d2).cuda(),
                      seed=torch.from_numpy(seed).cuda())
        din = rr["dL_din"].cpu().numpy()
        out.update({f"{name}_W": W, f"{name}_x": x, f"{name}_dL": dL, f"{name}_d2": d2, f"{name}_out": rr["out"].cpu().numpy(),
                    f"{name}_dL_din": din.T.copy() if soa else din, f"{name}_wgrad": rr["wgrad"].cpu().numpy()})
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_kernels.npz"))
