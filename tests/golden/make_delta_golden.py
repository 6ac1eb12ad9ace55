"""Generates tests/golden/delta_reference.npz by running the REFERENCE's own rotation-6d / DeltaNetwork kernels
(include/neural-graphics-primitives/common_operation.cuh, compiled as host code from where they lie into
oracle/_ref/libref_delta_host.so by oracle/Makefile.ref) on seeded inputs.  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_delta_golden.py tests/golden/delta_reference.npz

The committed .npz lets the test-suite pin the oracle's restatement (oracle/neus_oracle.cpp: delta_apply, delta_backward_sample,
orc_accumulate_global_movement) against reference outputs where the reference tree is not mounted (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main(out_path):
    from oracle import ref_delta_host as RD
    rs = np.random.RandomState(777)
    out = {}
    cases = []
    for k, (rs_sigma, ts_sigma, off) in enumerate([(0.0, 0.0, None), (0.03, 0.02, None), (0.2, 0.1, None), (0.05, 0.03, (0.5, 0.4, 0.6))]):
        rot6 = (np.array([1, 0, 0, 0, 1, 0], np.float32) + rs.normal(0, rs_sigma, 6)).astype(np.float16).astype(np.float32)
        trans = rs.normal(0, ts_sigma, 3).astype(np.float16).astype(np.float32)
        n = 512
        coords = rs.uniform(0, 1, (n, 7)).astype(np.float32)
        dL = (rs.normal(0, 1, (n, 3)) * 10.0 ** rs.uniform(-5, -1, (n, 1))).astype(np.float32)     # magnitudes down into the half subnormals
        offa = None if off is None else np.asarray(off, np.float32)
        pos, dirs = RD.apply(coords, rot6, trans, offa)
        each, dinput = RD.backward_each(coords, dL, rot6, trans, offa)
        acc_R = np.eye(3, dtype=np.float32)
        acc_t = np.zeros(3, np.float32)
        chain = []
        for _ in range(3):
            acc_R, acc_t = RD.accumulate(rot6, trans, acc_R, acc_t)
            chain.append(np.concatenate([acc_R.reshape(9), acc_t]))
        saved_R, saved_t = RD.accumulate(rot6, trans, chain[0][:9].reshape(3, 3), chain[0][9:], save_only=True)
        pre = f"c{k}_"
        out.update({pre + "rot6": rot6, pre + "trans": trans, pre + "off": np.zeros(3, np.float32) if offa is None else offa, pre + "coords": coords, pre + "dL": dL,
                    pre + "R": RD.rotation_matrix(rot6), pre + "pos": pos, pre + "dirs": dirs, pre + "each": each, pre + "dinput": dinput,
                    pre + "chain": np.stack(chain), pre + "saved": np.concatenate([saved_R.reshape(9), saved_t])})
        cases.append(k)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items() if k.startswith("c1_")})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "delta_reference.npz"))
