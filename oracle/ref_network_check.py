"""Checks the product's fused network kernels (through the C ABI) and the CPU oracle against the REFERENCE's own NerfNetwork<__half>
(oracle/ref_network.py) on the same seeded half parameters and inputs.  TEST INFRASTRUCTURE ONLY; run on a GPU box as

    python -m oracle.ref_network_check [--n 4096]

Torch-free (ctypes + libcudart), so it starts in a second.  The last stdout line is a JSON object {"ref_network_check": {...}, "ok": bool};
exit code 0 iff every bound holds.  tests/test_ref_network.py runs it in a subprocess (a fault inside the reference library cannot take the
test session down).  Bounds: the reference accumulates the MLP dot products in fp16 inside HMMA and sums grid gradients with fp16 atomics,
the product and the oracle in fp32 -- same bounds as tests/test_ref_kernels.py uses for the tcnn layer alone.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROT6 = np.array([1, 0.02, -0.01, 0.015, 1, 0.03], np.float32)
TRANS = np.array([0.01, -0.004, 0.002], np.float32)


def _h(bits):
    return bits.view(np.float16).astype(np.float32)


def _fwd_metrics(a_bits, b_bits):
    from tests import util as U
    a, b = _h(a_bits[:, :11]), _h(b_bits[:, :11])
    rel = np.abs(a - b) / (np.abs(b) + 1e-2)
    # per output column, relative to the column's scale (BASELINE.json states its fp16 tolerance that way): rgb 0..2, sdf 3, normal 4..6
    col = (np.abs(a - b).mean(axis=0) / (np.abs(b).mean(axis=0) + 1e-30))[:7]
    return dict(ulp_p99=float(np.percentile(U.half_ulp_diff(a_bits[:, :11], b_bits[:, :11]), 99)), mean_rel=float(rel.mean()), median_rel=float(np.median(rel)),
                max_rel=float(rel.max()), mean_abs_over_column_scale=[float(x) for x in col], copies_equal=bool(np.array_equal(a_bits[:, 7:11], b_bits[:, 7:11])))


def _seg_rel(x, y, lay):
    out = {}
    for name, a, b in [("sdf", 0, lay["rgb"]), ("rgb", lay["rgb"], lay["grid"]), ("grid", lay["grid"], lay["variance"])]:
        u, v = x[a:b].astype(np.float64), y[a:b].astype(np.float64)
        out[name] = float(np.linalg.norm(u - v) / (np.linalg.norm(v) + 1e-30))
    v = lay["variance"]
    out["variance"] = float(abs(x[v] - y[v]) / (abs(y[v]) + 1e-12))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096, help="samples (multiple of 128)")
    ap.add_argument("--no-oracle", action="store_true", help="skip the CPU oracle legs")
    ap.add_argument("--sdf", action="store_true", help="also compare NerfNetwork::sdf / ::density (lattice and occupancy-grid queries)")
    args = ap.parse_args()
    n = args.n
    assert n % 128 == 0
    import neus2_b200 as N
    from oracle import oracle as O
    from oracle import ref_network as RN
    from tests import util as U

    cuda = RN.Cuda()
    cfg = U.base_cfg()
    lay = O.param_layout(cfg)
    B = 1 << 14
    ctx = N.Context(N.default_config(batch_size=B))
    p = U.random_params(cfg, seed=1)
    ctx.set_params(p)
    lib = ctx._lib
    m, ph, e, g = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    N.check(lib.neus2_buffers(ctx._h, C.byref(m), C.byref(ph), C.byref(e), C.byref(g)))
    ref = RN.RefNetwork(per_level_scale=cfg.per_level_scale_py)
    res = {"n": n, "n_params_ref": ref.n_params, "n_params": int(ctx.n_params)}
    ok = ref.n_params == ctx.n_params
    fails = []

    def bound(name, value, limit):
        nonlocal ok
        if not (value <= limit):
            ok = False
            fails.append(f"{name}={value:.4g} > {limit}")

    if ok:
        cuda.sync()
        ref.set_params(ph.value, ctx.n_params)       # the product's own fp16 parameter buffer, device to device
        p_bits = cuda.download(ph.value, (ctx.n_params,), np.uint16)
        coords = U.random_coords(n, seed=11)
        rs = np.random.RandomState(12)
        dL = np.zeros((n, 16), dtype=np.float16)
        dL[:, :11] = (rs.normal(0, 1, (n, 11)) * np.array([1e-2] * 3 + [1e-1] + [30.0] * 3 + [1e-2] + [1e-2] * 3)).astype(np.float16)
        d_coords, d_dL = cuda.upload(coords), cuda.upload(dL)
        d_out_p, d_out_r = cuda.zeros(n * 32), cuda.zeros(n * 32)
        d_gr = cuda.zeros(ctx.n_params * 2)
        d_nin = cuda.upload(np.array([n], dtype=np.uint32))

        def product_forward():
            N.check(lib.neus2_forward(ctx._h, None, C.c_void_p(d_coords), C.c_uint32(n), C.c_void_p(d_out_p), C.c_int(0)))
            cuda.sync()
            return cuda.download(d_out_p, (n, 16), np.uint16)

        def product_backward():
            N.check(lib.neus2_backward(ctx._h, None, C.c_void_p(d_coords), C.c_void_p(d_dL), C.c_void_p(d_nin), C.c_uint32(n), C.c_uint32(B)))
            cuda.sync()
            return cuda.download(g.value, (ctx.n_params,), np.float32)

        # ---- forward, all levels (step 0) and progressive levels (step 150 -> 4 of 14) ----------------------------------------
        for step, valid in ((0, cfg.n_levels), (150, 4)):
            ctx.set_training_step(step)
            ref.set_training_step(step)
            out_p = product_forward()
            ref.forward(d_coords, n, d_out_r)
            out_r = cuda.download(d_out_r, (n, 16), np.uint16)
            r = {"product_vs_reference": _fwd_metrics(out_p, out_r)}
            if not args.no_oracle:
                out_o = O.net_forward(U.base_cfg(valid_level=valid), coords, p_bits)
                r["oracle_vs_reference"] = _fwd_metrics(out_o, out_r)
            res[f"forward_step{step}"] = r
            for who, mm in r.items():
                bound(f"forward_step{step}.{who}.mean_rel", mm["mean_rel"], 5e-3)
                bound(f"forward_step{step}.{who}.copies_differ", 0.0 if mm["copies_equal"] else 1.0, 0.0)
        ctx.set_training_step(0)
        ref.set_training_step(0)

        # ---- backward: parameter gradients (first + second order) ------------------------------------------------------------
        g_p = product_backward()
        ref.backward(d_coords, d_dL, n, B, d_out_r, d_gr)
        g_r = cuda.download(d_gr, (ctx.n_params,), np.float16).astype(np.float32)
        r = {"product_vs_reference": _seg_rel(g_p, g_r, lay)}
        if not args.no_oracle:
            g_o = O.net_backward(cfg, coords, p_bits, dL.view(np.uint16), B)
            r["oracle_vs_reference"] = _seg_rel(g_o, g_r, lay)
        res["backward"] = r
        for who, mm in r.items():
            for seg, v in mm.items():
                bound(f"backward.{who}.{seg}", v, 5e-2)

        # ---- NerfNetwork::sdf (marching-cubes lattice) and ::density (occupancy-grid update) -----------------------------------------
        if args.sdf:
            pos = np.random.RandomState(5).uniform(0, 1, (n, 3)).astype(np.float32)
            d_pos, d_sp, d_sr = cuda.upload(pos), cuda.zeros(n * 2), cuda.zeros(n * 2)
            for as_density in (False, True):
                N.check(lib.neus2_sdf(ctx._h, None, C.c_void_p(d_pos), C.c_uint32(3), C.c_uint32(n), C.c_void_p(d_sp), C.c_int(int(as_density)), C.c_int(0)))
                cuda.sync()
                ref.sdf(d_pos, n, d_sr, as_density)
                a, b = cuda.download(d_sp, (n,), np.uint16), cuda.download(d_sr, (n,), np.uint16)
                legs = {"product_vs_reference": (a, b)}
                if not args.no_oracle:
                    legs["oracle_vs_reference"] = (O.net_sdf(cfg, pos, p_bits, as_density=as_density), b)
                key = "density" if as_density else "sdf"
                res[key] = {}
                for who, (x, y) in legs.items():
                    xf, yf = _h(x), _h(y)
                    mm = dict(ulp_p99=float(np.percentile(U.half_ulp_diff(x, y), 99)), mean_abs_over_scale=float(np.abs(xf - yf).mean() / (np.abs(yf).mean() + 1e-30)))
                    res[key][who] = mm
                    bound(f"{key}.{who}.mean_abs_over_scale", mm["mean_abs_over_scale"], 5e-3)

        # ---- DeltaNetwork: deformation in the forward pass, 9 gradient sums in the backward pass ------------------------------
        try:
            ctx.set_global_movement(delta_rotation_6d=ROT6, delta_transition=TRANS, stream=0)
            ctx.set_delta_mode(True)
            ref.set_delta(True, True, ROT6, TRANS)
            acc = C.c_void_p()
            N.check(lib.neus2_delta_gradients_device(ctx._h, C.byref(acc)))
            out_p = product_forward()
            ref.forward(d_coords, n, d_out_r)
            out_r = cuda.download(d_out_r, (n, 16), np.uint16)
            fm = _fwd_metrics(out_p, out_r)
            cuda.memset0(acc.value, 48)
            g_p = product_backward()
            got = cuda.download(acc.value, (12,), np.float32)
            dref = ref.backward(d_coords, d_dL, n, B, d_out_r, d_gr, want_delta=True)
            g_r = cuda.download(d_gr, (ctx.n_params,), np.float16).astype(np.float32)
            want = np.concatenate([dref[:3], dref[4:10]])
            dm = dict(product=got[:9].tolist(), reference=want.tolist(), max_abs_over_scale=float(np.abs(got[:9] - want).max() / (np.abs(want).max() + 1e-30)))
            res["delta"] = {"forward": fm, "params": _seg_rel(g_p, g_r, lay), "delta_gradients": dm}
            bound("delta.forward.mean_rel", fm["mean_rel"], 5e-3)
            bound("delta.forward.copies_differ", 0.0 if fm["copies_equal"] else 1.0, 0.0)
            for seg, v in res["delta"]["params"].items():
                bound(f"delta.params.{seg}", v, 5e-2)
            bound("delta.delta_gradients", dm["max_abs_over_scale"], 5e-2)
        finally:
            ctx.set_delta_mode(False)
            ctx.set_global_movement(delta_rotation_6d=[1, 0, 0, 0, 1, 0], delta_transition=[0, 0, 0], stream=0)
    else:
        fails.append("parameter counts differ")
    ref.close()
    ctx.close()
    cuda.free_all()
    sys.stdout.flush()
    print(json.dumps({"ref_network_check": res, "fails": fails, "ok": bool(ok)}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
