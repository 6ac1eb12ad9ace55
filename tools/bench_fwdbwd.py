#!/usr/bin/env python
"""Times the training kernel (fused second forward + backward, k_network_fwdbwd) alone, on a FROZEN steady-state training state:
the DTU-shaped bench scene is trained for --preroll steps, the compacted batch of the last step (coordinates + dL/doutput) is
copied, and the kernel is launched on it repeatedly through the operator entry point (neus2_backward) with each profiling bit
of neus2_set_debug_bits.  CUDA-event stage timers of the library give the kernel time (and the weight-gradient GEMM's, when
NEUS2_WGRAD=staged).  Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preroll", type=int, default=700)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--bits", type=int, nargs="*", default=[0, 1, 4, 8, 16, 24, 25, 29])
    args = ap.parse_args()
    import torch
    import neus2_b200 as N
    from neus2_b200 import synthetic, _lib
    data = synthetic.dtu_like_views(n_views=49, width=1600, height=1200, seed=0)
    ctx = N.Context(N.default_config(), device=0)
    ctx.init_params()
    ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    for _ in range(args.preroll):
        st = ctx.train_step()
    torch.cuda.synchronize()
    B = 1 << 18
    coords = ctx.debug_buffer("coords_compacted", "<f4", (B, 7)).clone()
    dloss = ctx.debug_buffer("dloss", "<f2", (B, 16)).clone()
    n_in = int(st.measured_batch_size)
    lib = _lib.load()
    out = {"n_in": n_in, "rows": B, "wgrad": os.environ.get("NEUS2_WGRAD", "in-kernel"), "variants": {}}
    for bits in args.bits:
        lib.neus2_set_debug_bits(bits)
        ctx.set_profiling(True)
        for _ in range(args.reps):
            ctx.backward(coords, dloss, min(n_in, B), B, B)
        torch.cuda.synchronize()
        t = ctx.stage_times()
        out["variants"][str(bits)] = {"fwdbwd_ms": round(t["fwdbwd"][0], 4), "wgrad_ms": round(t["wgrad"][0], 4)}
        ctx.set_profiling(False)
    lib.neus2_set_debug_bits(0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
