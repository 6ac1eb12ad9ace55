"""Pins the CPU oracle (and through it the product) against the REFERENCE's own kernels: dependencies/my_tcnn compiled
for sm_100 from the reference sources in place (oracle/Makefile.ref -> oracle/_ref/libtcnn_ref.so).

tcnn layer only (hash grid fwd / bwd / 2nd order, FullyFusedMLP fwd / bwd / 2nd order); the reference's own NerfNetwork on
top of them is compared in tests/test_ref_network.py.  Tolerances: the reference accumulates MLP dot products in fp16 inside HMMA and grid gradients with
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


# ---- more of the reference's own code behind the oracle: optimizer, RNG, Morton codes -------------------------------------------
def test_reference_pcg32_and_morton_match_oracle():
    """pcg32 (tcnn/dependencies/pcg32/pcg32.h) and morton3D (tcnn common_device.h:339-362) are Eigen-free headers: the oracle's
    restatements are compared with the reference's own code, bit for bit."""
    from oracle import ref_kernels as R
    for seed, adv in ((1337, 0), (1337, 8 * 12345), (42, (1 << 32) + 17), (7, 3)):
        u, f = R.pcg32_sequence(seed, adv, 64)
        a = O.Pcg32(seed=seed); a.advance(adv)
        b = a.copy()
        assert [a.next_uint() for _ in range(64)] == [int(x) for x in u]
        assert np.array_equal(np.array([b.next_float() for _ in range(64)], dtype=np.float32).view(np.uint32), f.view(np.uint32))
    rs = np.random.RandomState(0)
    for x, y, z in rs.randint(0, 1024, (200, 3)):
        m = R.morton3D(int(x), int(y), int(z))
        assert m == O.lib().orc_morton3D(int(x), int(y), int(z))
        assert (R.morton3D_invert(m), R.morton3D_invert(m >> 1), R.morton3D_invert(m >> 2)) == (x, y, z)
        assert R.morton3D_invert(m) == O.lib().orc_morton3D_invert(m)


def test_reference_generate_random_uniform_matches_oracle():
    """The hash-grid initialisation (tcnn random.h:67-99 via grid.h:2375-2380): values and the generator state afterwards."""
    from oracle import ref_kernels as R
    n = 100003
    ref, state = R.generate_random_uniform(1337, n, -1e-4, 1e-4)
    rng = O.Pcg32(seed=1337)
    got = O.generate_random_uniform(rng, n, -1e-4, 1e-4)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert (int(rng.st[0]), int(rng.st[1])) == state


def test_reference_adam_ema_matches_oracle():
    """Trainer::optimizer_step with base.json's Ema{ExponentialDecay{Adam}}: matrix weights (L2, always stepped) followed by
    hash-grid style weights (skipped when the gradient is exactly zero, private step counters), five steps with a moving sparsity
    pattern.  The reference takes fp16 gradients; they are exactly representable, so the oracle's fp32 input is the same number."""
    import torch
    from oracle import ref_kernels as R
    n, n_matrix = 40000, 11264
    rs = np.random.RandomState(3)
    master0 = rs.normal(0, 0.1, n).astype(np.float32)
    opt = R.RefOptimizer(n, n_matrix)
    master = torch.from_numpy(master0.copy()).cuda()
    params = master.half()
    mm = master0.copy(); ph = master0.astype(np.float16).view(np.uint16).copy()
    m1 = np.zeros(n, np.float32); m2 = np.zeros(n, np.float32); steps = np.zeros(n, np.uint32); ema = np.zeros(n, np.uint16)
    try:
        for step in range(1, 6):
            g16 = np.zeros(n, dtype=np.float16)
            touched = rs.choice(n - n_matrix, size=6000, replace=False) + n_matrix
            g16[touched] = rs.normal(0, 2.0, touched.size).astype(np.float16)
            g16[:n_matrix] = rs.normal(0, 2.0, n_matrix).astype(np.float16)
            ref_ema = opt.step(128.0, master, params, torch.from_numpy(g16).cuda())
            O.adam_ema_step(n_matrix, 128.0, 1e-3, 0.9, 0.99, 1e-15, 1e-6, 0.95, step, g16.astype(np.float32), mm, ph, m1, m2, steps, ema)
            # the oracle writes the reference's FMAs out explicitly; what is left is powf (glibc here, libdevice there) in the bias
            # correction: at most an ulp of the update, i.e. far below an ulp of most weights
            got = master.cpu().numpy()
            assert np.mean(got.view(np.uint32) != mm.view(np.uint32)) < 2e-2, (step, np.mean(got != mm))
            assert np.max(np.abs(got - mm)) <= 2.0 ** -24 * np.max(np.abs(mm)), (step, np.max(np.abs(got - mm)))
            assert np.mean(params.cpu().numpy().view(np.uint16) != ph) < 1e-3, step
            d = U.half_ulp_diff(ref_ema.cpu().numpy().view(np.uint16), ema)
            assert d.max() <= 1 and np.mean(d > 0) < 5e-3, (step, d.max(), np.mean(d > 0))
            # untouched hash-grid style weights did not move and kept their step counter
            idle = np.ones(n, bool); idle[:n_matrix] = False; idle[touched] = False
            assert np.array_equal(master.cpu().numpy()[idle & (steps == 0)], master0[idle & (steps == 0)])
    finally:
        opt.close()


def test_product_grid_initialisation_follows_the_reference_stream():
    """neus2_init_params: Trainer rng (std::seed_seq{1337} -> pcg32, trainer.h:56-59), Xavier draws for the two MLPs (11,264 floats,
    nerf_network.h:819-870), then the hash grid from generate_random_uniform(-1e-4, 1e-4) -- the same values as the reference kernel."""
    import neus2_b200 as N
    from oracle import ref_kernels as R
    ctx = N.Context(N.default_config(batch_size=1 << 14))
    try:
        ctx.init_params()
        p = ctx.get_params()
        g0, g1 = ctx.layout["grid"], ctx.layout["variance"]
        rng = O.Pcg32(seed=int(O.lib().orc_trainer_seed_word(1337)))
        rng.advance(g0)                                    # one draw per MLP weight
        want = O.generate_random_uniform(rng, g1 - g0, -1e-4, 1e-4)
        assert np.array_equal(p[g0:g1].view(np.uint32), want.view(np.uint32))
        # and the oracle's generator equals the reference kernel on that very stream
        ref, _ = R.generate_random_uniform(1337, 4096, -1e-4, 1e-4)
        chk = O.generate_random_uniform(O.Pcg32(seed=1337), 4096, -1e-4, 1e-4)
        assert np.array_equal(ref.view(np.uint32), chk.view(np.uint32))
        assert np.all(p[g1:] == np.float32(0.3))
    finally:
        ctx.close()


def test_product_adam_ema_is_bit_identical_to_the_reference_optimizer():
    """The PRODUCT's fused Adam+EMA kernel against the reference's own Ema{ExponentialDecay{Adam}} objects on the full 10.56 M
    parameter vector (11,264 matrix weights, then the hash grid), four steps with a moving sparse gradient: fp32 master weights,
    fp16 weights and the EMA copy must agree bit for bit."""
    import torch
    import neus2_b200 as N
    from oracle import ref_kernels as R
    ctx = N.Context(N.default_config(batch_size=1 << 14))
    opt = None
    try:
        ctx.init_params()
        master, params, ema, grads = ctx.buffers()
        n, n_matrix = ctx.n_params, ctx.layout["grid"]
        opt = R.RefOptimizer(n, n_matrix)
        r_master = master.clone()
        r_params = params.clone()
        gen = torch.Generator(device="cuda").manual_seed(7)
        for step in range(1, 5):
            g16 = torch.zeros(n, dtype=torch.float16, device="cuda")
            idx = torch.randint(n_matrix, n, (300000,), device="cuda", generator=gen)
            g16[idx] = torch.randn(idx.numel(), device="cuda", generator=gen).half() * 3
            g16[:n_matrix] = torch.randn(n_matrix, device="cuda", generator=gen).half()
            grads.copy_(g16.float())
            ctx.optimizer_step(128.0)
            r_ema = opt.step(128.0, r_master, r_params, g16)
            torch.cuda.synchronize()
            assert torch.equal(master.view(torch.int32), r_master.view(torch.int32)), (step, (master != r_master).sum().item())
            assert torch.equal(params.view(torch.int16), r_params.view(torch.int16)), step
            assert torch.equal(ema.view(torch.int16), r_ema.view(torch.int16)), (step, (ema != r_ema).sum().item())
    finally:
        if opt is not None:
            opt.close()
        ctx.close()


def test_reference_fill_rollover_matches_oracle():
    """a8: the batch is padded to exactly B rows by modular copy and the dL/doutput copies are rescaled by n_in / B
    (tcnn common_device.h:514-534, launched at testbed_nerf.cu:3922-3930)."""
    import torch
    from oracle import ref_kernels as R
    rs = np.random.RandomState(9)
    n, n_in = 4096, 1537
    dl = (rs.normal(0, 1, (n, 16)) * (rs.rand(n, 16) < 0.7)).astype(np.float16)
    co = rs.rand(n, 7).astype(np.float32)
    dl_t, co_t = torch.from_numpy(dl.copy()).cuda(), torch.from_numpy(co.copy()).cuda()
    R.fill_rollover(n, n_in, dl_t, co_t)
    dl_o, co_o = dl.view(np.uint16).copy(), co.copy()
    O.fill_rollover(n, n_in, dl_o, co_o)
    assert np.array_equal(dl_t.cpu().numpy().view(np.uint16), dl_o)
    assert np.array_equal(co_t.cpu().numpy().view(np.uint32), co_o.view(np.uint32))
    # nothing to do when the batch is already full or empty
    for k in (0, n):
        d2 = torch.from_numpy(dl.copy()).cuda()
        R.fill_rollover(n, k, d2, None)
        assert np.array_equal(d2.cpu().numpy().view(np.uint16), dl.view(np.uint16))
