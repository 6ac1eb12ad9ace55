"""Times the network part of one training step in the REFERENCE's own ngp::NerfNetwork<__half> (oracle/_ref/libngp_network_ref.so: nerf_network.h
on the reference's tiny-cuda-nn, compiled for sm_100 from the reference sources) on this GPU, at the shapes bench.py's step has:
an inference pass over N_u generated samples (src/testbed_nerf.cu:3854) and forward + backward over B compacted samples (:3941-3946).
Sample generation, the loss kernel, the ~dozen glue kernels of train_nerf_step and the optimizer are NOT included, so the sum is a LOWER bound
of the reference's step time.  Supersedes profiles/bench_ref_kernels.py (which summed the tcnn objects one by one).  Torch-free.

    gpurun -- 'python profiles/bench_ref_network.py [N_u] [B] > gpurun_out/ref_network_ms.json'
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import oracle as O
    from oracle import ref_network as RN
    from tests import util as U
    n_u = (int(sys.argv[1]) if len(sys.argv) > 1 else 632064) // 128 * 128
    B = (int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18) // 128 * 128
    cfg = U.base_cfg()
    cuda = RN.Cuda()
    ref = RN.RefNetwork(per_level_scale=cfg.per_level_scale_py)
    assert ref.n_params == O.param_layout(cfg)["total"]
    ref.set_params(cuda.upload(U.random_params(cfg, seed=1).astype(np.float16)), ref.n_params)
    ref.set_training_step(2000)          # all 14 levels
    n = max(n_u, B)
    rs = np.random.RandomState(0)
    dL = np.zeros((n, 16), dtype=np.float16)
    dL[:, :11] = (rs.normal(0, 1, (n, 11)) * np.array([1e-2] * 3 + [1e-1] + [30.0] * 3 + [1e-2] + [1e-2] * 3)).astype(np.float16)
    d_coords, d_dL = cuda.upload(U.random_coords(n, seed=1)), cuda.upload(dL)
    ms_f, ms_b = ref.time_step(d_coords, d_dL, n_u, B, B, warmup=3, iters=20)
    print(json.dumps({"reference_network": "ngp::NerfNetwork<__half> (nerf_network.h) on dependencies/my_tcnn, sm_100", "n_inference": n_u, "n_train": B,
                      "inference_ms": round(ms_f, 4), "forward_backward_ms": round(ms_b, 4), "network_part_of_step_ms": round(ms_f + ms_b, 4)}))
    ref.close()
    cuda.free_all()


if __name__ == "__main__":
    main()
