"""CPU pin of the oracle's occupancy-grid update against golden vectors produced by the REFERENCE's own kernels
(src/testbed_nerf.cu: generate_grid_samples_nerf_nonuniform, splat_grid_samples_nerf_max_nearest_neighbor, ema_grid_samples_nerf,
grid_to_bitfield, bitfield_max_pool) compiled as host code from where they lie (oracle/Makefile.ref,
tests/golden/make_density_grid_golden.py).  Everything is bit-exact: cell indices, jittered positions, the decayed grid, the
occupancy bitfield and its seven max-pooled mips."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_nerf_host as RK

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "density_grid_reference.npz")


def _mod():
    spec = importlib.util.spec_from_file_location("make_density_grid_golden", os.path.join(HERE, "golden", "make_density_grid_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_occupancy_grid_update_matches_reference_kernels():
    g, mod = np.load(GOLD), _mod()
    grid, dens = mod.inputs()
    for k in range(2):
        pos, idx = O.generate_grid_samples(mod.N, O.Pcg32(state=g["rng_state"]), int(g[f"s{k}_step"]), grid, 1, float(g[f"s{k}_thresh"]))
        assert np.array_equal(idx, g[f"s{k}_idx"])
        assert np.array_equal(pos.view(np.uint32), g[f"s{k}_pos"].view(np.uint32))
    assert not np.array_equal(g["s0_idx"], g["s1_idx"])
    work = grid.copy()
    O.splat_and_decay(np.ascontiguousarray(g["s1_idx"]), dens, 0.95, work)
    assert np.array_equal(work[g["grid_probe_idx"]].view(np.uint32), g["grid_probe"].view(np.uint32))
    assert np.array_equal(mod.digest(work), g["grid_sha"])
    assert (work[grid < 0] < 0).all()                                       # untrained cells stay marked
    bf, mean = O.update_bitfield(work, 1)
    assert abs(mean - float(g["mean"])) <= 1e-6 * float(g["mean"])          # the reference reduces in fp32 with an unspecified tree order
    pop = np.array([int(np.unpackbits(bf[mod.G3 // 8 * l: mod.G3 // 8 * (l + 1)]).sum()) for l in range(8)], np.int64)
    assert np.array_equal(pop, g["bitfield_popcount"]) and pop[0] > pop[1] > pop[2] > 0
    assert np.array_equal(mod.digest(bf), g["bitfield_sha"])


@pytest.mark.skipif(not RK.available(), reason="needs the reference tree (oracle/_ref/libref_nerf_host.so)")
def test_committed_fixture_is_what_the_reference_source_produces(tmp_path):
    g, mod = np.load(GOLD), _mod()
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(fresh[name], g[name]), name
