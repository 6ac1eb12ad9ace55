"""CPU pin of the oracle's low-discrepancy sequences and pcg32 against golden vectors produced by the REFERENCE's own source:
tests/golden/ld_reference.npz is frozen from include/neural-graphics-primitives/random_val.cuh (sobol, Laine-Karras nested uniform
scrambling, ld_random_val, ld_random_pixel_offset) and dependencies/my_tcnn/dependencies/pcg32/pcg32.h compiled as host code from where
they lie (oracle/Makefile.ref, tests/golden/make_ld_golden.py).  Integer work and the float conversions are bit-exact."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ld_reference.npz")
REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_random_host.so")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_sobol_and_ld_random_val_bit_exact(gold):
    idx, seeds = gold["idx"], gold["seeds"]
    sob = np.array([[O.sobol(int(i), d) for d in (0, 1)] for i in idx], dtype=np.uint32)
    assert np.array_equal(sob, gold["sobol"])
    ld = np.array([[O.ld_random_val(int(i), int(s)) for s in seeds] for i in idx], dtype=np.float32)
    assert np.array_equal(ld.view(np.uint32), gold["ld_random_val"].view(np.uint32))
    assert ld.min() >= 0.0 and ld.max() <= 1.0


def test_pixel_offsets_bit_exact(gold):
    off = np.array([O.ld_random_pixel_offset(int(s)) for s in gold["spps"]], dtype=np.float32)
    assert np.array_equal(off.view(np.uint32), gold["pixel_offset"].view(np.uint32))
    assert np.array_equal(off[0], [0.5, 0.5])                      # spp 0 (snap_to_pixel_centers) is the pixel centre


def test_pcg32_streams_bit_exact(gold):
    for seed, adv, f, u in zip(gold["pcg_seed"], gold["pcg_advance"], gold["pcg_floats"], gold["pcg_uints"]):
        r = O.Pcg32(seed=int(seed))
        r.advance(int(adv))
        r2 = r.copy()
        ff = np.array([r.next_float() for _ in range(len(f))], np.float32)
        uu = np.array([r2.next_uint() for _ in range(len(u))], np.uint32)
        assert np.array_equal(ff.view(np.uint32), f.view(np.uint32)) and np.array_equal(uu, u), (seed, adv)


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference/include")), reason="needs the reference tree (oracle/_ref/libref_random_host.so)")
def test_committed_fixture_is_what_the_reference_source_produces(gold, tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ld_golden", os.path.join(os.path.dirname(GOLD), "make_ld_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh_path = str(tmp_path / "fresh.npz")
    mod.main(fresh_path)
    fresh = np.load(fresh_path)
    assert sorted(fresh.files) == sorted(gold.files)
    for name in gold.files:
        assert np.array_equal(fresh[name], gold[name]), name
