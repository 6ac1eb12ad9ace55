"""Generates tests/golden/density_grid_reference.npz by running the REFERENCE's own occupancy-grid kernels (src/testbed_nerf.cu:
generate_grid_samples_nerf_nonuniform :640-669, splat_grid_samples_nerf_max_nearest_neighbor :671-690, ema_grid_samples_nerf :710-740,
grid_to_bitfield :748-772, bitfield_max_pool :774-795) compiled as host code from where they lie (oracle/Makefile.ref ->
oracle/_ref/libref_nerf_host.so).  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_density_grid_golden.py tests/golden/density_grid_reference.npz

Inputs are rebuilt from seeds by the test (functions below); large outputs are stored as SHA-256 digests plus per-level popcounts.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

G3 = 128 ** 3
N = 4096


def inputs():
    rs = np.random.RandomState(5)
    grid = np.where(rs.rand(G3) < 0.02, rs.rand(G3) * 3, 0).astype(np.float32)
    grid[rs.randint(0, G3, 1000)] = -1.0                       # cells marked untrained stay negative through the EMA
    dens = (rs.rand(N) * 2).astype(np.float16).view(np.uint16)
    return grid, dens


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def main(out_path):
    from oracle import oracle as O
    from oracle import ref_nerf_host as RK
    grid, dens = inputs()
    r = O.Pcg32(seed=42); r.advance()
    st = r.st.copy()
    out = {"rng_state": st}
    for k, (step, thresh) in enumerate([(3, -0.01), (7, 0.01)]):   # uniform phase (thresh < 0: first candidate always taken) and the thresholded search
        pos, idx = RK.grid_samples(N, st, step, grid, 1, thresh)
        out[f"s{k}_step"], out[f"s{k}_thresh"], out[f"s{k}_pos"], out[f"s{k}_idx"] = np.array(step), np.float32(thresh), pos, idx
    g = grid.copy()
    RK.splat_ema(out["s1_idx"], dens, 0.95, g)
    out["grid_sha"] = digest(g)
    out["grid_probe_idx"] = np.unique(np.concatenate([out["s1_idx"][:256], np.arange(0, G3, 65537, dtype=np.uint32)]))
    out["grid_probe"] = g[out["grid_probe_idx"]]
    mean = np.float32(np.maximum(g, 0).astype(np.float64).sum() / G3)
    bf = RK.bitfield(g, 1, float(mean))
    out["mean"] = mean
    out["bitfield_sha"] = digest(bf)
    out["bitfield_popcount"] = np.array([int(np.unpackbits(bf[G3 // 8 * l: G3 // 8 * (l + 1)]).sum()) for l in range(8)], np.int64)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes", out["bitfield_popcount"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "density_grid_reference.npz"))
