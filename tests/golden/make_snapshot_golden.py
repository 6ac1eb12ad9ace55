"""Generates tests/golden/snapshot_reference.msgpack (+ .npz): a snapshot file WRITTEN BY THE REFERENCE'S OWN CODE -- its Ema{ExponentialDecay{Adam}}
optimizer classes stepped on the CPU and serialised by their own serialize(), its GPUMemory -> json binary conversion and its json::to_msgpack, compiled
as host code from where they lie (oracle/Makefile.ref -> oracle/_ref/libref_snapshot_host.so; Trainer::serialize and Testbed::save_snapshot, a dozen lines
each, are restated around them in oracle/ref_snapshot_host.cpp).  CPU only; needs /root/reference:

    make -f oracle/Makefile.ref all && python tests/golden/make_snapshot_golden.py

A miniature NerfNetwork-shaped parameter vector (two small MLPs, a 'hash grid', 4 variance slots) takes STEPS optimizer steps on seeded half gradients
(sparse on the grid, like the real thing); the .npz holds the inputs and the optimizer state the reference ended in.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

LAYERS = [(16, 8), (4, 16), (16, 12), (4, 16)]          # density MLP: 16x8, 4x16; rgb MLP: 16x12, 4x16
N_DENSITY, N_RGB, N_GRID, N_VARIANCE = 16 * 8 + 4 * 16, 16 * 12 + 4 * 16, 600, 4
STEPS = 45
LOSS_SCALE = 128.0
OPTIMIZER = {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 20, "decay_interval": 10, "decay_base": 0.33,
             "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}}
NETWORK_CONFIG = {"encoding": {"otype": "HashGrid", "n_levels": 14}, "optimizer": OPTIMIZER, "loss": {"otype": "Huber"}}
TESTBED = dict(training_step=STEPS, loss=0.0123, rays_per_batch=4096, measured=262000, measured_before=631000, aabb_scale=1,
               rotation=[1, 0, 0, 0, 1, 0, 0, 0, 1], transition=[0.25, -0.5, 0.125], local_rotation=[1, 0.0625, 0, 0, 1, -0.03125], local_transition=[0.5, 0, -0.25])


def inputs():
    n = N_DENSITY + N_RGB + N_GRID + N_VARIANCE
    rs = np.random.RandomState(11)
    master = rs.normal(0, 0.1, n).astype(np.float32)
    grads = np.zeros((STEPS, n), np.float16)
    for s in range(STEPS):
        grads[s, :N_DENSITY + N_RGB] = rs.normal(0, 2.0, N_DENSITY + N_RGB)
        touched = rs.choice(N_GRID, size=150, replace=False) + N_DENSITY + N_RGB
        grads[s, touched] = rs.normal(0, 2.0, touched.size)
        grads[s, n - N_VARIANCE] = rs.normal(0, 2.0)                              # only the first variance slot ever gets a gradient
    return master, grads


def run_reference():
    from oracle import ref_snapshot_host as RS
    master, grads = inputs()
    st = RS.Stack(OPTIMIZER, N_DENSITY, N_RGB, N_GRID, N_VARIANCE, LAYERS)
    st.set_params_full_precision(master)
    for s in range(STEPS):
        st.optimizer_step(LOSS_SCALE, grads[s])
    t = TESTBED
    st.set_testbed_state(t["training_step"], t["loss"], t["rays_per_batch"], t["measured"], t["measured_before"], t["aabb_scale"], None, t["rotation"], t["transition"],
                         t["local_rotation"], t["local_transition"])
    return st


def main(out_dir):
    st = run_reference()
    path = os.path.join(out_dir, "snapshot_reference.msgpack")
    st.save_snapshot(NETWORK_CONFIG, path, True)
    s = st.state()
    np.savez_compressed(os.path.join(out_dir, "snapshot_reference.npz"), **{k: np.asarray(v) for k, v in s.items()})
    st.close()
    print("wrote", path, os.path.getsize(path), "bytes; optimizer step", s["current_step"], "lr", s["learning_rate"], "x", s["learning_rate_factor"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
