"""Distribution of samples per ray (marched / kept by the loss) in the steady state of the bench workload -- what the tail of the per-ray
kernels (loss, first pass) is made of.  gpurun -- 'python tools/ray_length_probe.py > gpurun_out/ray_lengths.json'"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    import neus2_b200 as N
    args = argparse.Namespace(views=49, width=1600, height=1200)
    data = bench.make_data(args)
    B = 1 << 18
    ctx = N.Context(N.default_config(batch_size=B), device=0)
    ctx.init_params()
    ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    for _ in range(705):
        ctx.train_step()
    out = []
    for _ in range(3):
        st0 = ctx.get_state()
        ctx.train_step_begin()
        torch.cuda.synchronize()
        cnt = ctx.debug_buffer("counters", "<u4", (8,)).cpu().numpy()
        R, k = st0["rays_per_batch"], int(cnt[0])
        idx = ctx.debug_buffer("ray_indices", "<u4", (R,)).cpu().numpy()[:k]
        marched = ctx.debug_buffer("numsteps_all", "<u4", (R,)).cpu().numpy()[idx]
        kept = ctx.debug_buffer("numsteps", "<u4", (R, 2)).cpu().numpy()[:k, 0]
        q = [50, 90, 99, 99.9, 100]
        out.append({"rays": R, "kept_rays": k, "samples": int(cnt[1]), "compacted": int(cnt[2]), "first_pass_evaluated": int(cnt[5]),
                    "marched_per_ray_percentiles": dict(zip(map(str, q), np.percentile(marched, q).tolist())),
                    "kept_per_ray_percentiles": dict(zip(map(str, q), np.percentile(kept, q).tolist())),
                    "rays_with_more_than_32_64_128_256_kept": [int((kept > t).sum()) for t in (32, 64, 128, 256)],
                    "rays_with_more_than_32_64_128_256_marched": [int((marched > t).sum()) for t in (32, 64, 128, 256)]})
        ctx.train_step_end()
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
