"""CPU-only pin of the oracle against golden vectors produced by the REFERENCE's own kernels.

tests/golden/ref_kernels.npz was generated on a B200 by tests/golden/make_golden.py from dependencies/my_tcnn compiled in
place (oracle/Makefile.ref): hash-grid forward / backward / second-order backward and the two fused MLPs
(forward / backward / second-order backward).  The reference repository itself ships no golden vectors.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import util as U

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _h(a):
    return a.view(np.float16).astype(np.float32) if a.dtype == np.uint16 else a.astype(np.float32)


def _grid(gold, cfg):
    lay = O.param_layout(cfg)
    grid = np.zeros(lay["n_grid"], dtype=np.float16)
    grid[gold["grid_touched"]] = gold["grid_values"]
    return grid.view(np.uint16)


def test_grid_tables_match_reference_param_count(gold):
    cfg = U.base_cfg()
    assert abs(cfg.per_level_scale_py - float(gold["per_level_scale"])) == 0
    assert O.param_layout(cfg)["n_grid"] == 10548128            # reference GridEncoding::n_params() for base.json


def test_grid_forward_bit_exact_vs_reference(gold):
    cfg = U.base_cfg()
    enc, _ = O.grid_forward(cfg, np.ascontiguousarray(gold["grid_coords"]), _grid(gold, cfg))
    assert np.array_equal(enc, gold["grid_enc"].view(np.uint16))       # hash indices + half accumulation order: bit exact


def test_grid_progressive_mask_vs_reference(gold):
    cfg = U.base_cfg(valid_level=4)      # set_training_step(150): ceil(0.2*14 + 0.02*50) = 4, levels 0..4 live
    enc, _ = O.grid_forward(cfg, np.ascontiguousarray(gold["grid_coords"]), _grid(gold, cfg))
    assert np.array_equal(enc, gold["grid_enc_step150"].view(np.uint16))
    assert not enc[10:].any() and enc[:10].any()


def test_grid_backward_and_second_order_vs_reference(gold):
    cfg = U.base_cfg()
    coords = np.ascontiguousarray(gold["grid_coords"])
    dL_dy = np.ascontiguousarray(gold["grid_dL_dy"]).view(np.uint16)
    v = np.ascontiguousarray(gold["grid_v"])
    _, dydx = O.grid_forward(cfg, coords, _grid(gold, cfg))
    assert np.allclose(O.grid_backward_input(cfg, dL_dy, dydx), gold["grid_dL_dx"], rtol=1e-4, atol=1e-6)
    ddy = O.grid_bwdbwd_dLdoutput(cfg, v, dydx)
    assert U.half_ulp_diff(ddy, gold["grid_dL_ddLdy"].view(np.uint16)).max() <= 1
    gg = np.zeros(O.param_layout(cfg)["n_grid"], dtype=np.float32)
    O.grid_backward(cfg, coords, dL_dy, gg)
    O.grid_bwdbwd_grid(cfg, coords, v, dL_dy, gg)
    t = gold["grid_touched"]
    ref = gold["grid_grad"].astype(np.float32)
    assert np.linalg.norm(gg[t] - ref) / np.linalg.norm(ref) < 5e-3     # reference: fp16 atomics, oracle: fp32 sums
    untouched = np.ones(gg.shape[0], bool)
    untouched[t] = False
    assert np.abs(gg[untouched]).max() < 1e-6                           # nothing outside the reference's touched set (up to fp16 underflow)


@pytest.mark.parametrize("name,n_in,hl", [("sdf", 32, 1), ("rgb", 48, 2)])
def test_mlp_vs_reference(gold, name, n_in, hl):
    W = gold[f"{name}_W"].view(np.uint16)
    x = np.ascontiguousarray(gold[f"{name}_x"]).view(np.uint16)
    dL = np.ascontiguousarray(gold[f"{name}_dL"]).view(np.uint16)
    d2 = np.ascontiguousarray(gold[f"{name}_d2"]).view(np.uint16)
    n = x.shape[0]
    seed = np.zeros((n, 16), dtype=np.float16)
    seed[:, 0] = 1
    hid, out = O.mlp_forward(W, x, n_in, hl)
    ref_out = gold[f"{name}_out"].astype(np.float32)
    assert np.max(np.abs(_h(out) - ref_out)) < 4e-3 * np.abs(ref_out).max()      # reference accumulates in fp16 inside HMMA
    wg = np.zeros(W.shape[0], dtype=np.float32)
    din = O.mlp_backward(W, x, hid, dL, n_in, hl, wgrad=wg)
    O.mlp_bwdbwd(W, hid, d2, seed.view(np.uint16), n_in, hl, wg)
    ref_din = gold[f"{name}_dL_din"].astype(np.float32)
    assert np.linalg.norm(_h(din) - ref_din) / np.linalg.norm(ref_din) < 3e-2
    ref_wg = gold[f"{name}_wgrad"].astype(np.float32)
    assert np.linalg.norm(wg - ref_wg) / np.linalg.norm(ref_wg) < 3e-2
