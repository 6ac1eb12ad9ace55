"""GPU parity of one full training step, stage by stage, against the CPU oracle.

Each stage of neus2_train_step_begin is checked in isolation by feeding the ORACLE the GPU's output of the previous
stage (so a half-ulp difference upstream cannot masquerade as, or hide, a difference downstream):
occupancy-grid samples -> density -> grid/bitfield -> ray marching -> network -> loss/compaction -> backward.
Integer work (cell indices, bitfield, per-ray sample counts, offsets) must match bit-exactly.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu

B = 1 << 14
G3 = 128 ** 3


@pytest.fixture(scope="module", params=[1, 0], ids=["early_termination", "full_forward"])
def stepped(request):
    import torch
    import neus2_b200 as N
    from neus2_b200 import synthetic
    cfg = U.base_cfg()
    ctx = N.Context(N.default_config(batch_size=B, mask_loss_weight=0.1, forward_early_termination=request.param))
    p = U.random_params(cfg, seed=2)
    # make the random network look like an SDF of a blob so that rays find occupied cells: bias the sdf output
    ctx.set_params(p)
    data = synthetic.render_sphere_views(n_views=5, width=96, height=80, seed=1)
    ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    # step 0 runs in full (all levels on: grid.h:2429-2432); the staged comparison is done on step 1, where the
    # progressive mask is active (valid_level = ceil(0.2*14) = 3) and the occupancy grid is non-trivial.
    ctx.set_params(p)
    st_first = ctx.train_step()
    torch.cuda.synchronize()
    grid_before = ctx.debug_buffer("density_grid", "<f4", (G3,)).cpu().numpy().copy()
    p1 = ctx.get_params()
    state0 = ctx.get_state()
    ctx.train_step_begin()
    torch.cuda.synchronize()
    yield dict(ctx=ctx, cfg=cfg, p=p1, data=data, state0=state0, st_first=st_first, grid_before=grid_before, early=bool(request.param))
    ctx.close()


def _np(t):
    return t.cpu().numpy()


def _step_params(stepped):
    """(rays_per_batch, max_inference, n_rays_total) of the staged step, from the first step's stats (testbed_nerf.cu:3770-3790)."""
    st = stepped["st_first"]
    R = stepped["state0"]["rays_per_batch"]
    assert R == st.rays_per_batch_next
    max_inf = (min(st.measured_batch_size_before_compaction, B * 16) + 127) // 128 * 128
    return R, max_inf, st.rays_per_batch


def test_density_grid_stage(stepped):
    ctx, cfg, p = stepped["ctx"], stepped["cfg"], stepped["p"]
    # density_grid_rng = pcg32(m_rng.next_uint()), m_rng = pcg32(seed)   (testbed.cu:2087-2100)
    m_rng = O.Pcg32(seed=1337)
    dg_rng = O.Pcg32(seed=m_rng.next_uint())
    dg_rng.advance(); dg_rng.advance()            # the step-0 update advanced it twice (testbed_nerf.cu:3349,3362)
    grid0 = stepped["grid_before"]
    pos, idx = O.generate_grid_samples(G3, dg_rng, 1, grid0, 1, -0.01)
    g_pos = _np(ctx.debug_buffer("grid_positions", "<f4", (G3, 3)))
    g_idx = _np(ctx.debug_buffer("grid_indices", "<u4", (G3,)))
    # the same samples, written in CELL order instead of sample order (density_grid.cu: the uniform samples' cell index is an invertible
    # map of the sample index; neighbours in space become neighbours in the buffer the network pass reads)
    order = np.argsort(idx, kind="stable")
    assert np.array_equal(np.sort(idx), np.unique(idx))                  # every cell at most once
    assert np.array_equal(g_idx, idx[order])                             # integer: bit exact
    assert np.array_equal(g_pos.view(np.uint32), pos[order].view(np.uint32))   # fp32 positions: bit exact (same op order)
    # density: network in half
    sub = np.arange(0, G3, 97)
    dens_ref = O.net_sdf(U.base_cfg(valid_level=3), np.ascontiguousarray(g_pos[sub]), U.to_half_bits(p), as_density=True)
    g_dens = _np(ctx.debug_buffer("grid_density", "<f2", (G3,))).view(np.uint16)
    d = U.half_ulp_diff(g_dens[sub], dens_ref)
    assert np.percentile(d, 99) <= 8, np.percentile(d, [50, 99, 100])
    # splat + decay + bitfield from the GPU's own densities
    grid = grid0.copy()
    O.splat_and_decay(g_idx, g_dens, 0.95, grid)
    g_grid = _np(ctx.debug_buffer("density_grid", "<f4", (G3,)))
    assert np.array_equal(g_grid.view(np.uint32), grid.view(np.uint32))
    bitfield, mean = O.update_bitfield(g_grid, 1)
    g_mean = float(_np(ctx.debug_buffer("density_grid_mean", "<f4", (1,)))[0])
    assert abs(g_mean - mean) <= 1e-5 * abs(mean)
    g_bits = _np(ctx.debug_buffer("bitfield", "<u1", (G3 // 8 * 8,)))
    if min(0.1, g_mean) == min(0.1, mean):
        assert np.array_equal(g_bits, bitfield)
    else:   # threshold differs in the last bit of an fp32 mean: at most a handful of cells may flip
        assert np.count_nonzero(np.unpackbits(g_bits ^ bitfield)) <= 8


def test_density_grid_samples_after_step_256():
    """From step 256 on an occupancy refresh draws 128^3 / 4 uniform + 128^3 / 4 non-uniform samples (testbed_nerf.cu:4011).  The uniform
    ones are written in cell order through a prefix sum across blocks (density_grid.cu); every sample -- cell index and position bits --
    is the reference's, the non-uniform ones keep their sample order."""
    import torch
    import neus2_b200 as N
    from neus2_b200 import synthetic
    ctx = N.Context(N.default_config(batch_size=B))
    try:
        ctx.init_params()
        data = synthetic.render_sphere_views(n_views=3, width=64, height=48, seed=1)
        ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
        ctx.set_state(training_step=300, rays_per_batch=4096)
        n = G3 // 4
        m_rng = O.Pcg32(seed=1337)
        dg_rng = O.Pcg32(seed=m_rng.next_uint())
        grid = np.zeros(G3, dtype=np.float32)
        for ema_step in range(2):                                          # the second refresh sees a non-trivial grid and another step offset
            ctx.update_density_grid()
            torch.cuda.synchronize()
            pos_u, idx_u = O.generate_grid_samples(n, dg_rng, ema_step, grid, 1, -0.01)
            dg_rng.advance()
            pos_n, idx_n = O.generate_grid_samples(n, dg_rng, ema_step, grid, 1, 0.1)
            dg_rng.advance()
            g_pos = _np(ctx.debug_buffer("grid_positions", "<f4", (G3, 3)))[:2 * n]
            g_idx = _np(ctx.debug_buffer("grid_indices", "<u4", (G3,)))[:2 * n]
            order = np.argsort(idx_u, kind="stable")
            assert len(np.unique(idx_u)) == n
            assert np.array_equal(g_idx[:n], idx_u[order]) and np.array_equal(g_pos[:n].view(np.uint32), pos_u[order].view(np.uint32))
            assert np.array_equal(g_idx[n:], idx_n) and np.array_equal(g_pos[n:].view(np.uint32), pos_n.view(np.uint32))
            g_dens = _np(ctx.debug_buffer("grid_density", "<f2", (G3,)))[:2 * n].view(np.uint16)
            O.splat_and_decay(g_idx, g_dens, 0.95, grid)
            assert np.array_equal(_np(ctx.debug_buffer("density_grid", "<f4", (G3,))).view(np.uint32), grid.view(np.uint32))
    finally:
        ctx.close()


def test_sample_generation_stage(stepped):
    ctx, cfg, data, st0 = stepped["ctx"], stepped["cfg"], stepped["data"], stepped["state0"]
    R, max_inf, n_rays_total = _step_params(stepped)
    bitfield = _np(ctx.debug_buffer("bitfield", "<u1", (G3 // 8 * 8,)))
    ds = O.make_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    rng = O.Pcg32(state=st0["rng_state"])
    ref = O.generate_training_samples(R, max_inf, n_rays_total, rng, ds, bitfield)
    cnt = _np(ctx.debug_buffer("counters", "<u4", (4,)))
    assert int(cnt[0]) == ref["n_kept"] and int(cnt[1]) == ref["n_samples"]
    assert ref["n_kept"] > 100 and ref["n_samples"] > 1000, "degenerate test scene"
    assert np.array_equal(_np(ctx.debug_buffer("numsteps_all", "<u4", (R,))), ref["numsteps_all"])       # per-ray sample counts: bit exact
    k = ref["n_kept"]
    assert np.array_equal(_np(ctx.debug_buffer("ray_indices", "<u4", (R,)))[:k], ref["ray_indices"][:k])
    # ("numsteps" was rewritten in place by the loss stage with the compacted counts; it is checked there)
    assert np.array_equal(_np(ctx.debug_buffer("rays", "<f4", (R, 6)))[:k].view(np.uint32), ref["rays"][:k].view(np.uint32))
    last = ref["numsteps"][k - 1]
    n = int(last[0] + last[1])                                                                            # samples of the kept rays
    g_coords = _np(ctx.debug_buffer("coords", "<f4", (B * 16, 7)))[:n]
    for i in range(k):                                                                                    # dropped rays leave holes
        c, b = ref["numsteps"][i]
        assert np.array_equal(g_coords[b:b + c].view(np.uint32), ref["coords"][b:b + c].view(np.uint32))  # sample positions: bit exact
    stepped["gen"] = ref


def test_forward_stage(stepped):
    ctx, cfg, p = stepped["ctx"], stepped["cfg"], stepped["p"]
    cnt = _np(ctx.debug_buffer("counters", "<u4", (4,)))
    R, max_inf, _ = _step_params(stepped)
    k = int(cnt[0])
    ns = _np(ctx.debug_buffer("numsteps_all", "<u4", (R,)))
    n = min(int(cnt[1]), max_inf)
    coords = _np(ctx.debug_buffer("coords", "<f4", (B * 16, 7)))[:n]
    sub = np.arange(0, min(n, int(ns[: R // 2].sum())), max(1, n // 4000))   # stay inside rows written by kept rays
    if stepped["early"]:
        # the first pass stops along each ray where the loss stops reading: the rows the loss consumed (the ray's first
        # `compacted` samples, testbed_nerf.cu:1567-1571) must all be there; rows behind the cut are never looked at
        g_ns = _np(ctx.debug_buffer("numsteps", "<u4", (R, 2)))[:k]          # rewritten by the loss: (compacted count, compacted base)
        idx = _np(ctx.debug_buffer("ray_indices", "<u4", (R,)))[:k]
        bases = np.concatenate([[0], np.cumsum(ns[idx].astype(np.int64))[:-1]]).astype(np.int64)
        rows = np.concatenate([np.arange(b, b + int(c), dtype=np.int64) for b, c in zip(bases, g_ns[:, 0]) if c > 0])
        assert len(rows) == int(cnt[2]) or int(cnt[2]) > B
        sub = rows[:: max(1, len(rows) // 4000)]
        sub = sub[sub < n]
    cfg0 = U.base_cfg(valid_level=3)       # training step 1: valid_level = ceil(0.2*14) = 3  (grid.h:2433)
    ref = O.net_forward(cfg0, np.ascontiguousarray(coords[sub]), U.to_half_bits(p))
    out = _np(ctx.debug_buffer("mlp_out", "<f2", (B * 16, 16)))[:n].view(np.uint16)[sub]
    d = U.half_ulp_diff(out[:, :11], ref[:, :11])
    assert np.percentile(d, 99) <= 2, np.percentile(d, [50, 99, 100])


def test_loss_stage(stepped):
    ctx, cfg, data, st0 = stepped["ctx"], stepped["cfg"], stepped["data"], stepped["state0"]
    R, max_inf, n_rays_total = _step_params(stepped)
    cnt = _np(ctx.debug_buffer("counters", "<u4", (4,)))
    k, n = int(cnt[0]), min(int(cnt[1]), max_inf)
    bitfield = _np(ctx.debug_buffer("bitfield", "<u1", (G3 // 8 * 8,)))
    ds = O.make_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    gen = O.generate_training_samples(R, max_inf, n_rays_total, O.Pcg32(state=st0["rng_state"]), ds, bitfield)
    mlp_out = _np(ctx.debug_buffer("mlp_out", "<f2", (B * 16, 16)))[:max_inf].view(np.uint16).copy()
    lp = O.LossParams()
    lp.loss_scale = 128.0
    lp.random_bg_color = 1
    lp.snap_to_pixel_centers = 1
    lp.mask_loss_weight = 0.1
    lp.ek_loss_weight = 0.01
    lp.cos_anneal_ratio = 1.0
    ref = O.compute_loss(R, k, n_rays_total, O.Pcg32(state=st0["rng_state"]), B, lp, ds, mlp_out, gen["ray_indices"], gen["rays"], gen["numsteps"], gen["coords"])
    assert int(cnt[2]) == ref["n_compacted"]                                                       # compacted sample count
    g_ns = _np(ctx.debug_buffer("numsteps", "<u4", (R, 2)))[:k]
    assert np.array_equal(g_ns, ref["numsteps"][:k])                                                # per-ray compacted counts + offsets: bit exact
    m = min(ref["n_compacted"], B)
    g_cc = _np(ctx.debug_buffer("coords_compacted", "<f4", (B, 7)))[:m]
    assert np.array_equal(g_cc.view(np.uint32), ref["coords_compacted"][:m].view(np.uint32))
    g_dl = _np(ctx.debug_buffer("dloss", "<f2", (B, 16)))[:m].astype(np.float32)
    r_dl = ref["dloss"][:m].view(np.float16).astype(np.float32)
    # fp32 math with expf/powf/logf from different libms (CUDA vs glibc): 1e-5 relative, then one rounding to half
    err = np.abs(g_dl[:, :11] - r_dl[:, :11])
    tol = 2e-3 * np.abs(r_dl[:, :11]) + 1e-6 * np.max(np.abs(r_dl[:, :11]), axis=0, keepdims=True)
    assert np.mean(err > tol) < 1e-3, np.mean(err > tol)
    for name, rtol in [("loss", 2e-5), ("ek_loss", 2e-5), ("mask_loss", 5e-3)]:   # mask: log(1-w) amplifies 1e-7 errors of w near 1
        g = _np(ctx.debug_buffer(name, "<f4", (R,)))[:k]
        assert np.allclose(g, ref[name][:k], rtol=rtol, atol=1e-9), name


def test_backward_stage(stepped):
    ctx, cfg, p = stepped["ctx"], stepped["cfg"], stepped["p"]
    cnt = _np(ctx.debug_buffer("counters", "<u4", (4,)))
    n_in = min(int(cnt[2]), B)
    coords = _np(ctx.debug_buffer("coords_compacted", "<f4", (B, 7))).copy()
    dl = _np(ctx.debug_buffer("dloss", "<f2", (B, 16))).view(np.uint16).copy()
    O.fill_rollover(B, n_in, dl_h=dl, coords=coords)
    cfg0 = U.base_cfg(valid_level=3)
    ref = O.net_backward(cfg0, coords, U.to_half_bits(p), dl, B)
    g = _np(ctx.buffers()[3])
    lay = O.param_layout(cfg)
    for name, a, b in [("sdf", 0, lay["rgb"]), ("rgb", lay["rgb"], lay["grid"]), ("grid", lay["grid"], lay["variance"])]:
        x, y = g[a:b].astype(np.float64), ref[a:b].astype(np.float64)
        rel = np.linalg.norm(x - y) / (np.linalg.norm(y) + 1e-30)
        assert rel < 3e-3, (name, rel)
    # levels above valid_level receive no gradient (grid.h:394, 906)
    off, _, _ = O.cfg_tables(cfg)
    assert not g[lay["grid"] + 2 * int(off[4]):lay["variance"]].any()


def test_step_end_and_training_progress(stepped):
    ctx = stepped["ctx"]
    R = stepped["state0"]["rays_per_batch"]
    st = ctx.train_step_end()
    assert st.training_step == 2 and st.loss_valid == 0
    assert stepped["st_first"].loss_valid == 1 and np.isfinite(stepped["st_first"].loss)     # refreshed every 16 steps (testbed.cu:2714)
    assert st.measured_batch_size > 0 and st.rays_per_batch_next % 128 == 0
    # rays_per_batch <- rays * target / measured, rounded up to 128, capped at 2^18   (testbed_nerf.cu:3434-3435)
    want = min((int(np.float32(np.float32(R) * np.float32(B) / np.float32(st.measured_batch_size))) + 127) // 128 * 128, 1 << 18)
    assert st.rays_per_batch_next == want
    losses = []
    for _ in range(48):
        s = ctx.train_step()
        if s.loss_valid:
            losses.append(s.loss)
    assert all(np.isfinite(losses))


@pytest.mark.parametrize("variant", ["early_termination", "fused_loss"])
def test_early_termination_changes_nothing_downstream(variant, monkeypatch):
    """Two contexts, same seed / data / parameters, first network pass with and without the early cut: every buffer the loss
    writes is bit-identical step after step (sample counts, compacted coordinates, dL/doutput, per-ray losses), and the first
    pass evaluates far fewer samples.  Variant "fused_loss": the same comparison between the loss by samples (default: k_loss_rays replays
    the recurrences over what the first pass stored, k_loss_samples does the per-sample work one thread per compacted sample) and the
    one-warp-per-ray composite / scan / gradient kernels (NEUS2_LOSS_BY_SAMPLES=0, read per step), incl. the eikonal / mask terms."""
    import torch
    import neus2_b200 as N
    from neus2_b200 import synthetic
    data = synthetic.render_sphere_views(n_views=6, width=128, height=96, seed=5)
    ctxs = []
    fused_env = ("1", "1") if variant == "early_termination" else ("1", "0")
    for et in ((1, 0) if variant == "early_termination" else (1, 1)):
        c = N.Context(N.default_config(batch_size=B, forward_early_termination=et, mask_loss_weight=0.1))
        c.init_params()
        p = c.get_params()
        p[c.layout["variance"]:] = 0.55          # a sharp surface: rays saturate early, the cut has something to cut
        c.set_params(p)
        c.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
        ctxs.append(c)
    try:
        evaluated = []
        longest = 0
        for step in range(6):
            outs = []
            for c, env in zip(ctxs, fused_env):
                monkeypatch.setenv("NEUS2_LOSS_BY_SAMPLES", env)
                c.train_step_begin()
                torch.cuda.synchronize()
                cnt = _np(c.debug_buffer("counters", "<u4", (8,))).copy()
                R = c.get_state()["rays_per_batch"]
                k = int(cnt[0])
                m = min(int(cnt[2]), B)
                outs.append(dict(cnt=cnt[:3].copy(), work=cnt[4:6].copy(), numsteps=_np(c.debug_buffer("numsteps", "<u4", (R, 2)))[:k].copy(),
                                 cc=_np(c.debug_buffer("coords_compacted", "<f4", (B, 7)))[:m].copy(), dl=_np(c.debug_buffer("dloss", "<f2", (B, 16)))[:m].view(np.uint16).copy(),
                                 loss=_np(c.debug_buffer("loss", "<f4", (R,))).copy(), ek=_np(c.debug_buffer("ek_loss", "<f4", (R,))).copy(),
                                 mask=_np(c.debug_buffer("mask_loss", "<f4", (R,))).copy()))
                # keep both replicas on the same trajectory: the early-cut context adopts nothing, both apply their own gradients;
                # gradients differ only by fp32 atomic order, so re-synchronise the parameters after the step
                c.train_step_end()
            a, b = outs
            assert np.array_equal(a["cnt"], b["cnt"]), (step, a["cnt"], b["cnt"])
            assert np.array_equal(a["numsteps"], b["numsteps"])
            assert np.array_equal(a["cc"].view(np.uint32), b["cc"].view(np.uint32))
            assert np.array_equal(a["dl"], b["dl"])
            assert np.array_equal(a["loss"].view(np.uint32), b["loss"].view(np.uint32))
            assert np.array_equal(a["ek"].view(np.uint32), b["ek"].view(np.uint32)) and np.array_equal(a["mask"].view(np.uint32), b["mask"].view(np.uint32))
            longest = max(longest, int(a["numsteps"][:, 0].max()))
            evaluated.append((int(a["work"][1]), int(a["cnt"][1])))
            ctxs[1].set_params(ctxs[0].get_params(), reset_optimizer=False)      # same fp32 master -> same fp16 parameters next step
            for t_dst, t_src in zip(ctxs[1].buffers()[1:3], ctxs[0].buffers()[1:3]):
                t_dst.copy_(t_src)
            m1a, m2a, sa, _ = ctxs[0].optimizer_buffers()
            m1b, m2b, sb, _ = ctxs[1].optimizer_buffers()
            m1b.copy_(m1a); m2b.copy_(m2a); sb.copy_(sa)
            ga = ctxs[0].debug_buffer("density_grid", "<f4", (G3,)); gb = ctxs[1].debug_buffer("density_grid", "<f4", (G3,))
            gb.copy_(ga)
            ctxs[1].set_bitfield(_np(ctxs[0].debug_buffer("bitfield", "|u1", (G3 // 8 * 8,))))
        assert longest > 32, longest              # rays longer than one 32-sample chunk of the loss kernels were covered
        ev, tot = evaluated[-1]
        assert 0 < ev <= tot
        if variant == "early_termination":
            assert ev < 0.9 * tot, (ev, tot)          # the cut removes work
    finally:
        for c in ctxs:
            c.close()


def test_pipelined_step_generates_the_same_samples():
    """config.prefetch_next_samples (default on): the samples of step k+1 are generated during step k on a second stream, from the
    step state the host already knows after step k's loss kernel.  They must be exactly what the in-order step generates: checked
    against the oracle on steps that used the prefetch, and the schedule of prefetched steps is the complement of the
    occupancy-refresh schedule (a refresh changes the bitfield the samples depend on)."""
    import torch
    import neus2_b200 as N
    from neus2_b200 import synthetic
    data = synthetic.render_sphere_views(n_views=6, width=128, height=96, seed=5)
    ds = O.make_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    c = N.Context(N.default_config(batch_size=B))
    assert c.cfg.prefetch_next_samples == 1
    try:
        c.init_params()
        c.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
        checked = 0
        for step in range(45):
            due = step % min(max(step // 16, 1), 16) == 0
            st0 = c.get_state()
            assert st0["training_step"] == step
            check = (not due) and step in (33, 37, 43)
            if check:
                bitfield = _np(c.debug_buffer("bitfield", "<u1", (G3 // 8 * 8,))).copy()
            c.train_step_begin()
            if check:
                torch.cuda.synchronize()
                R = st0["rays_per_batch"]
                max_inf = (min(st0["measured_before"], B * 16) + 127) // 128 * 128
                ref = O.generate_training_samples(R, max_inf, st0["n_rays_total"], O.Pcg32(state=st0["rng_state"]), ds, bitfield)
                cnt = _np(c.debug_buffer("counters", "<u4", (8,)))
                assert int(cnt[0]) == ref["n_kept"] and int(cnt[1]) == ref["n_samples"] and ref["n_kept"] > 20
                assert np.array_equal(_np(c.debug_buffer("numsteps_all", "<u4", (R,))), ref["numsteps_all"])
                k = ref["n_kept"]
                assert np.array_equal(_np(c.debug_buffer("ray_indices", "<u4", (R,)))[:k], ref["ray_indices"][:k])
                assert np.array_equal(_np(c.debug_buffer("rays", "<f4", (R, 6)))[:k].view(np.uint32), ref["rays"][:k].view(np.uint32))
                last = ref["numsteps"][k - 1]
                g_coords = _np(c.debug_buffer("coords", "<f4", (B * 16, 7)))[:int(last[0] + last[1])]
                for i in range(k):
                    n_i, b_i = ref["numsteps"][i]
                    assert np.array_equal(g_coords[b_i:b_i + n_i].view(np.uint32), ref["coords"][b_i:b_i + n_i].view(np.uint32))
                checked += 1
            st = c.train_step_end()
            assert st.measured_batch_size > 0
            assert st.samples_prefetched == (0 if (due or step == 0) else 1), (step, due, st.samples_prefetched)
            assert st.density_grid_updated == int(due)
        assert checked == 3
        # step 45's samples are in flight now; a call that changes what sample generation reads discards them
        c.set_state(**{k: c.get_state()[k] for k in ("training_step", "rays_per_batch", "n_rays_total", "measured_before")})
        flags = [c.train_step().samples_prefetched for _ in range(3)]       # 45: discarded, 46: occupancy refresh, 47: prefetched again
        assert flags == [0, 0, 1], flags
    finally:
        c.close()

