"""CPU checks of the oracle's DeltaNetwork restatement (include/neural-graphics-primitives/transform_network.h:23-480,
common_operation.cuh:37-160, 416-470, 551-580, 788-840): the 6d-rotation forward against a float64 reference, the hand-derived
backward (gradient_rotation_matrix_to_6d) against torch autograd, and the accumulation of frame-to-frame motions.  These pin the
MATH; the pin against the reference's own source is tests/test_golden_delta.py."""
import numpy as np
import pytest
import torch

from oracle import oracle as O


def _half(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def _fwd64(r6, t, x):
    a1, a2 = r6[:3], r6[3:]
    b1 = a1 / a1.norm()
    u = a2 - (b1 @ a2) * b1
    b2 = u / u.norm()
    R = torch.stack([b1, b2, torch.linalg.cross(b1, b2)], 1)
    return (x + t) @ R.T, R


@pytest.mark.parametrize("seed", [0, 1])
def test_delta_forward_and_backward_match_autograd(seed):
    rs = np.random.RandomState(seed)
    rot6 = _half(np.array([1, 0, 0, 0, 1, 0], np.float32) + rs.normal(0, 0.03, 6))
    trans = _half(rs.normal(0, 0.02, 3))
    n = 4096
    coords = rs.uniform(0, 1, (n, 7)).astype(np.float32)
    L = rs.normal(0, 1e-2, (n, 3)).astype(np.float32)
    r6 = torch.tensor(rot6, dtype=torch.float64, requires_grad=True)
    t = torch.tensor(trans, dtype=torch.float64, requires_grad=True)
    p, R = _fwd64(r6, t, torch.tensor(coords[:, :3], dtype=torch.float64))
    out = O.delta_apply(coords, rot6, trans)
    assert np.abs(out[:, :3] - p.detach().numpy()).max() < 5e-7
    assert np.array_equal(out[:, 3], coords[:, 3])                                  # dt passes through
    d = (torch.tensor(coords[:, 4:7] * 2 - 1, dtype=torch.float64) @ R.T).detach().numpy()
    assert np.abs(out[:, 4:7] - (d + 1) / 2).max() < 5e-7                          # directions: un-warp, rotate, warp
    Rm, tm = O.delta_matrix(rot6, trans)
    assert np.abs(Rm - R.detach().numpy()).max() < 2e-7 and np.array_equal(tm, trans)
    assert np.abs(Rm @ Rm.T - np.eye(3)).max() < 1e-6
    (p * torch.tensor(L, dtype=torch.float64)).sum().backward()
    g = O.delta_backward(coords, L, rot6, trans)
    # per-sample gradients are rounded to half before the fp32 batch sum (the reference's T = __half matrices): ~1e-3 relative
    ref = np.concatenate([t.grad.numpy(), [0], r6.grad.numpy(), [0, 0]])
    assert np.abs(g - ref).max() < 3e-3 * np.abs(ref).max() + 1e-4
    assert g[3] == 0 and g[10] == 0 and g[11] == 0


def test_identity_delta_is_a_no_op():
    coords = np.random.RandomState(3).uniform(0, 1, (257, 7)).astype(np.float32)
    out = O.delta_apply(coords, [1, 0, 0, 0, 1, 0], [0, 0, 0])
    assert np.array_equal(out[:, :4], coords[:, :4])
    assert np.abs(out[:, 4:] - coords[:, 4:]).max() < 1e-7                         # (2d - 1 + 1) / 2 rounds


def test_accumulate_composes_motions():
    rs = np.random.RandomState(5)
    R_acc, t_acc = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    R64, t64 = np.eye(3), np.zeros(3)
    for _ in range(4):
        rot6 = _half(np.array([1, 0, 0, 0, 1, 0], np.float32) + rs.normal(0, 0.05, 6))
        trans = _half(rs.normal(0, 0.02, 3))
        Rd, td = O.delta_matrix(rot6, trans)
        R_acc, t_acc = O.accumulate_global_movement(rot6, trans, R_acc, t_acc)
        R64, t64 = Rd.astype(np.float64) @ R64, Rd.astype(np.float64) @ (t64 + td)
        # stored in half after every frame (common_operation.cuh:567-579)
        assert np.abs(R_acc - R64).max() < 2e-3 and np.abs(t_acc - t64).max() < 2e-4
        assert np.array_equal(R_acc, _half(R_acc))
    # a ray moved by the accumulated motion == moved through each frame's delta in turn (up to the half storage)
    x = rs.uniform(0, 1, 3)
    assert np.abs((R_acc @ x + t_acc) - (R64 @ x + t64)).max() < 5e-3


def test_net_backward_delta_consistency():
    """backward with the delta == backward on the deformed coordinates; the delta gradient is the chain rule on its dL/dpos."""
    from tests import util as U
    cfg = U.base_cfg()
    p = U.to_half_bits(U.random_params(cfg, seed=2))
    n = 256
    coords = U.random_coords(n, seed=4)
    rs = np.random.RandomState(6)
    dL = np.zeros((n, 16), dtype=np.float16)
    dL[:, :11] = (rs.normal(0, 1, (n, 11)) * np.array([1e-2] * 3 + [1e-1] + [30.0] * 3 + [1e-2] + [1e-2] * 3)).astype(np.float16)
    rot6, trans = _half([1, 0.02, -0.01, 0.015, 1, 0.03]), _half([0.01, -0.004, 0.002])
    g, dg, dpos = O.net_backward_delta(cfg, coords, p, dL.view(np.uint16), n, rot6, trans)
    deformed = O.delta_apply(coords, rot6, trans)
    g2 = O.net_backward(cfg, deformed, p, dL.view(np.uint16), n)
    assert np.array_equal(g, g2)
    assert np.array_equal(dg, O.delta_backward(coords, dpos, rot6, trans))
    assert np.abs(dpos).max() > 0 and np.isfinite(dpos).all()
    # identity delta: canonical gradients equal the plain backward's bit for bit
    g3, _, _ = O.net_backward_delta(cfg, coords, p, dL.view(np.uint16), n, [1, 0, 0, 0, 1, 0], [0, 0, 0])
    g4 = O.net_backward(cfg, O.delta_apply(coords, [1, 0, 0, 0, 1, 0], [0, 0, 0]), p, dL.view(np.uint16), n)
    assert np.array_equal(g3, g4)
