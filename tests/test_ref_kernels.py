"""Pins the CPU oracle (and through it the product) against the REFERENCE's own kernels: dependencies/my_tcnn compiled
for sm_100 from the reference sources in place (oracle/Makefile.ref -> oracle/_ref/libtcnn_ref.so).

tcnn layer only (hash grid fwd / bwd / 2nd order, FullyFusedMLP fwd / bwd / 2nd order); the ngp layer does not compile
here.  Tolerances: the reference accumulates MLP dot products in fp16 inside HMMA and grid gradients with
fp16 atomics, the oracle in fp32 -- BASELINE.json's 1e-3 (relative to the tensor's scale) for fp16 paths applies.
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_kernels as RK
from tests import util as U

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RK.available(), reason="oracle/_ref/libtcnn_ref.so not built")]


def _h(a):
    return a.view(np.float16).astype(np.float32)


def test_reference_grid_forward_backward_matches_oracle():
    import torch
    cfg = U.base_cfg()
    lay = O.param_layout(cfg)
    p = U.random_params(cfg, seed=5)
    grid_h = U.to_half_bits(p[lay["grid"]:lay["variance"]])
    n = 4096
    coords = U.random_coords(n, seed=6)
    rs = np.random.RandomState(7)
    dL_dy = rs.normal(0, 1e-2, (28, n)).astype(np.float16)
    v = rs.normal(0, 1e-2, (n, 3)).astype(np.float32)
    g = RK.RefGrid(per_level_scale=cfg.per_level_scale_py)
    assert g.n_params == lay["n_grid"]
    r = g.eval(torch.from_numpy(coords).cuda(), torch.from_numpy(grid_h.view(np.float16)).cuda(), torch.from_numpy(dL_dy).cuda(),
               torch.from_numpy(v).cuda(), torch.from_numpy(dL_dy).cuda())
    enc, dydx = O.grid_forward(cfg, coords, grid_h)
    # encoding: same gathers, same half accumulation order -> bit exact (hash indices included)
    assert np.array_equal(r["enc"].cpu().numpy().view(np.uint16), enc)
    # dL/dx = sum_k dL/dy_k dy_k/dx in fp32
    dx = O.grid_backward_input(cfg, dL_dy.view(np.uint16), dydx)
    assert np.allclose(r["dL_dx"].cpu().numpy(), dx, rtol=1e-4, atol=1e-6)
    # second order -> dL/d(dL/dy)
    ddy = O.grid_bwdbwd_dLdoutput(cfg, v, dydx)
    assert np.percentile(U.half_ulp_diff(r["dL_ddLdy"].cpu().numpy().view(np.uint16), ddy), 99.9) <= 1
    # grid gradient: first order (Overwrite) + second order (Accumulate); reference sums with fp16 atomics
    gg = np.zeros(lay["n_grid"], dtype=np.float32)
    O.grid_backward(cfg, coords, dL_dy.view(np.uint16), gg)
    O.grid_bwdbwd_grid(cfg, coords, v, dL_dy.view(np.uint16), gg)
    rg = r["grad"].cpu().numpy().astype(np.float32)
    assert np.array_equal(rg != 0, gg != 0) or np.mean((rg != 0) != (gg != 0)) < 1e-3
    rel = np.linalg.norm(rg - gg) / np.linalg.norm(gg)
    assert rel < 5e-3, rel


@pytest.mark.parametrize("n_in,hl,soa", [(32, 1, True), (48, 2, False)])
def test_reference_mlp_matches_oracle(n_in, hl, soa):
    import torch
    rs = np.random.RandomState(n_in)
    n = 2048
    n_params = 64 * n_in + (hl - 1) * 4096 + 1024
    W = rs.uniform(-0.25, 0.25, n_params).astype(np.float16)
    x = rs.uniform(-1, 1, (n, n_in)).astype(np.float16)
    dL = np.zeros((n, 16), dtype=np.float16)
    dL[:, :4] = rs.normal(0, 1e-2, (n, 4)).astype(np.float16)
    d2 = rs.normal(0, 1e-2, (n, n_in)).astype(np.float16)
    seed = np.zeros((n, 16), dtype=np.float16)
    seed[:, 0] = 1
    m = RK.RefMlp(n_in, hl)
    assert m.n_params == n_params
    xin = torch.from_numpy(np.ascontiguousarray(x.T) if soa else x).cuda()
    d2in = torch.from_numpy(d2).cuda()   # always [n][in]: see ref_shim.cu
    r = m.eval(xin, torch.from_numpy(W).cuda(), in_soa=soa, dL_dout=torch.from_numpy(dL).cuda(), d_in2=d2in, seed=torch.from_numpy(seed).cuda())
    hid, out = O.mlp_forward(W.view(np.uint16), x.view(np.uint16), n_in, hl)
    ro = r["out"].cpu().numpy().astype(np.float32)
    scale = np.abs(_h(out)).max()
    assert np.max(np.abs(ro - _h(out))) < 4e-3 * scale
    wg = np.zeros(n_params, dtype=np.float32)
    din = O.mlp_backward(W.view(np.uint16), x.view(np.uint16), hid, dL.view(np.uint16), n_in, hl, wgrad=wg)
    O.mlp_bwdbwd(W.view(np.uint16), hid, d2.view(np.uint16), seed.view(np.uint16), n_in, hl, wg)
    rd = r["dL_din"].cpu().numpy().astype(np.float32)
    rd = rd.T if soa else rd
    assert np.linalg.norm(rd - _h(din)) / np.linalg.norm(_h(din)) < 3e-2   # fp16 HMMA accumulation in the reference
    rw = r["wgrad"].cpu().numpy().astype(np.float32)
    assert np.linalg.norm(rw - wg) / np.linalg.norm(wg) < 3e-2
