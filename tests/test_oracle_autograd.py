"""Oracle self-check: the restated NerfNetwork backward (first + second order) must equal the true gradient.

The reference's backward (nerf_network.h:330-601) is, for ReLU MLPs and linear grid interpolation, the exact
gradient of  L = <dL/dout, out>  w.r.t. all parameters (the omitted terms only reach positions).
So a float64 autograd model of the same network pins the oracle's semantics (signs, slots, normalisers);
differences are fp16 rounding only.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O


def _sh4_torch(dw):
    x, y, z = dw[:, 0] * 2 - 1, dw[:, 1] * 2 - 1, dw[:, 2] * 2 - 1
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [0.28209479177387814 + 0 * x, -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
         1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
         -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
         0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
         0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
         0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(o, dim=1)


def _torch_model(cfg, coords, params, dL_dout, B, masks):
    lay = O.param_layout(cfg)
    SI = lay["sdf_in"]
    p = torch.tensor(params.view(np.float16).astype(np.float64), requires_grad=True)
    W0 = p[0:64 * SI].view(64, SI)
    W1 = p[64 * SI:64 * SI + 1024].view(16, 64)
    o = lay["rgb"]
    C0 = p[o:o + 64 * 48].view(64, 48)
    C1 = p[o + 3072:o + 3072 + 4096].view(64, 64)
    C2 = p[o + 7168:o + 7168 + 1024].view(16, 64)
    grid = p[lay["grid"]:lay["grid"] + lay["n_grid"]].view(-1, 2)
    var = p[lay["variance"]]
    n = coords.shape[0]
    # the SDF-MLP sees half(x)-0.5 but the grid sees the fp32 x; derivative of both w.r.t. x is taken as identity
    x32 = torch.tensor(coords[:, :3].astype(np.float64), requires_grad=True)
    xh = torch.tensor(coords[:, :3].astype(np.float16).astype(np.float64))
    off, res, sc = O.cfg_tables(cfg)
    encs = []
    for l in range(cfg.n_levels):
        scale = float(sc[l])
        hs = int(off[l + 1] - off[l])
        pf = (coords[:, :3] * np.float32(scale) + np.float32(0.5)).astype(np.float32)
        cell = np.floor(pf).astype(np.int64)
        frac = x32 * scale + 0.5 - torch.tensor(cell.astype(np.float64))
        acc = 0
        for c in range(8):
            w = 1
            cc = cell.copy()
            for d in range(3):
                if c & (1 << d):
                    w = w * frac[:, d]
                    cc[:, d] += 1
                else:
                    w = w * (1 - frac[:, d])
            idx = np.array([O.grid_index(hs, int(res[l]), int(a), int(b), int(cz)) for a, b, cz in cc], dtype=np.int64) // 2
            acc = acc + w[:, None] * grid[int(off[l]) + torch.tensor(idx)]
        encs.append(acc)
    enc = torch.cat(encs, dim=1)
    pad = torch.zeros(n, SI - 3 - enc.shape[1], dtype=torch.float64)
    sdf_in = torch.cat([xh + (x32 - x32.detach()) - 0.5, enc, pad], dim=1)
    h0 = (sdf_in @ W0.T) * masks[0]
    so = h0 @ W1.T
    normal = torch.autograd.grad(so[:, 0].sum(), x32, create_graph=True)[0]
    sh = _sh4_torch(torch.tensor(coords[:, 4:7].astype(np.float64)))
    rgb_in = torch.cat([so, sh, xh, normal, torch.zeros(n, 10, dtype=torch.float64)], dim=1)
    out = ((((rgb_in @ C0.T) * masks[1]) @ C1.T) * masks[2]) @ C2.T
    g = torch.tensor(dL_dout.view(np.float16).astype(np.float64))
    L = (g[:, :3] * out[:, :3]).sum() + (g[:, 3] * so[:, 0]).sum() + (g[:, 4:7] * normal).sum() / B \
        + (g[:, 8:11] * normal).sum() + (g[:, 7]).sum() * var
    L.backward()
    return p.grad.numpy()


@pytest.mark.parametrize("mode,tol", [("sdf", 3e-3), ("ek", 3e-3), ("cos", 3e-3), ("rgb", 3e-3), ("all", 3e-3)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_net_backward_matches_autograd(seed, mode, tol):
    rs = np.random.RandomState(seed)
    cfg = O.make_cfg(n_levels=4, base_resolution=4, log2_hashmap_size=8, per_level_scale=1.7)
    lay = O.param_layout(cfg)
    n = 96
    params = np.zeros(lay["total"], dtype=np.float16)
    params[: lay["grid"]] = rs.uniform(-0.4, 0.4, lay["grid"]).astype(np.float16)
    params[lay["grid"]: lay["variance"]] = rs.uniform(-0.5, 0.5, lay["n_grid"]).astype(np.float16)
    params[lay["variance"]] = 0.3
    coords = rs.uniform(0.05, 0.95, (n, 7)).astype(np.float32)
    full = rs.normal(0, 1.0, (n, 11)).astype(np.float16)
    dL = np.zeros((n, 16), dtype=np.float16)
    B = 64
    if mode == "all":
        dL[:, :11] = full
    elif mode == "rgb":
        dL[:, :3] = full[:, :3]
    elif mode == "sdf":
        dL[:, 3] = full[:, 3]
    elif mode == "ek":
        dL[:, 4:7] = full[:, 4:7] * B
    elif mode == "cos":
        dL[:, 8:11] = full[:, 8:11]
    g_or = O.net_backward(cfg, coords, params.view(np.uint16), dL.view(np.uint16), B)
    # ReLU states are taken from the oracle's own forward so that a unit whose pre-activation rounds across zero in
    # fp16 (a real, measured effect: 2 flips of 6144 move the gradient by percents at n=96) cannot blur the comparison.
    _, dbg = O.net_forward(cfg, coords, params.view(np.uint16), debug=True)
    masks = [torch.tensor((dbg["sdf_hidden"][0].view(np.float16) > 0).astype(np.float64)),
             torch.tensor((dbg["rgb_hidden"][0].view(np.float16) > 0).astype(np.float64)),
             torch.tensor((dbg["rgb_hidden"][1].view(np.float16) > 0).astype(np.float64))]
    g_t = _torch_model(cfg, coords, params.view(np.uint16), dL.view(np.uint16), B, masks)
    for name, a, b in [("sdf", 0, lay["rgb"]), ("rgb", lay["rgb"], lay["grid"]), ("grid", lay["grid"], lay["variance"])]:
        x, y = g_or[a:b].astype(np.float64), g_t[a:b]
        if np.linalg.norm(y) == 0:
            assert np.linalg.norm(x) == 0, name
            continue
        rel = np.linalg.norm(x - y) / np.linalg.norm(y)
        cos = float(x @ y / (np.linalg.norm(x) * np.linalg.norm(y)))
        assert rel < tol and cos > 0.98, (name, rel, cos)
    if mode == "all":
        v = lay["variance"]
        assert abs(g_or[v] - g_t[v]) <= 2e-3 * abs(g_t[v]) + 1e-3
