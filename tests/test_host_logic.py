"""CPU-only tests: published known-answer vectors, host-side logic, and the C-ABI surface (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pcg32_published_known_answers():
    # pcg32-demo of the PCG reference implementation (pcg-c-basic): seed(42, 54) -> first six outputs.
    lib = O.lib()
    st = np.zeros(2, dtype=np.uint64)
    # Pcg32(initstate=42, initseq=54): orc_pcg32_seed uses initseq=1, so build the state by hand as pcg32.h:55-62 does
    inc = np.uint64((54 << 1) | 1)
    st[0], st[1] = 0, inc
    lib.orc_pcg32_next_uint(st.ctypes.data_as(O.u64p))
    st[0] = np.uint64((int(st[0]) + 42) & 0xFFFFFFFFFFFFFFFF)
    lib.orc_pcg32_next_uint(st.ctypes.data_as(O.u64p))
    got = [int(lib.orc_pcg32_next_uint(st.ctypes.data_as(O.u64p))) for _ in range(6)]
    assert got == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]


def test_pcg32_advance_equals_stepping():
    a, b = O.Pcg32(seed=1337), O.Pcg32(seed=1337)
    for _ in range(1000):
        a.next_uint()
    b.advance(1000)
    assert a.next_uint() == b.next_uint()
    f = O.Pcg32(seed=7).next_float()
    assert 0.0 <= f < 1.0


def test_morton_roundtrip_and_known_values():
    lib = O.lib()
    assert lib.orc_morton3D(1, 0, 0) == 1 and lib.orc_morton3D(0, 1, 0) == 2 and lib.orc_morton3D(0, 0, 1) == 4
    assert lib.orc_morton3D(127, 127, 127) == 128 ** 3 - 1
    for x, y, z in [(5, 77, 120), (127, 0, 64)]:
        m = lib.orc_morton3D(x, y, z)
        assert (lib.orc_morton3D_invert(m), lib.orc_morton3D_invert(m >> 1), lib.orc_morton3D_invert(m >> 2)) == (x, y, z)


def test_grid_tables_base_json():
    # SURVEY §8 table (float32 arithmetic of testbed.cu:2181 and grid.h:1472-1508)
    cfg = O.make_cfg()
    off, res, sc = O.cfg_tables(cfg)
    assert list(res) == [16, 24, 34, 50, 72, 104, 151, 219, 317, 461, 669, 971, 1411, 2049]
    assert list(sc) == [r - 1 for r in res]
    assert list(np.diff(off)) == [4096, 13824, 39304, 125000, 373248] + [524288] * 9
    lay = O.param_layout(cfg)
    assert lay["total"] == 10559396 and lay["grid"] == 11264 and lay["sdf_in"] == 32


def test_grid_index_dense_and_hash():
    # dense: x + y*res + z*res^2 (mod size); hashed: x ^ y*2654435761 ^ z*805459861 (mod 2^19)   (grid.h:119-153)
    assert O.grid_index(4096, 16, 3, 2, 1) == 2 * (3 + 2 * 16 + 1 * 256)
    assert O.grid_index(4096, 16, 16, 15, 15) == 2 * ((16 + 15 * 16 + 15 * 256) % 4096)      # +1 corner wraps
    x, y, z = 1000, 2000, 1234
    h = (x ^ (y * 2654435761 & 0xFFFFFFFF) ^ (z * 805459861 & 0xFFFFFFFF)) % (1 << 19)
    assert O.grid_index(1 << 19, 2049, x, y, z) == 2 * h


def test_srgb_transfer_functions():
    assert abs(O.lib().orc_srgb_to_linear(0.04045) - 0.04045 / 12.92) < 1e-9
    assert abs(O.lib().orc_linear_to_srgb(1.0) - 1.0) < 1e-6
    assert abs(O.lib().orc_srgb_to_linear(O.lib().orc_linear_to_srgb(0.5)) - 0.5) < 2e-4   # the reference's 0.41666 exponent is not exactly 1/2.4


def test_fill_rollover_semantics():
    # rows < n_in untouched; wrapped copies of dL are rescaled by n_in/n (tcnn common_device.h:515-534)
    n, n_in = 16, 5
    dl = np.arange(n * 16, dtype=np.float32).reshape(n, 16).astype(np.float16)
    co = np.arange(n * 7, dtype=np.float32).reshape(n, 7)
    d2, c2 = dl.view(np.uint16).copy(), co.copy()
    O.fill_rollover(n, n_in, dl_h=d2, coords=c2)
    d2 = d2.view(np.float16)
    assert np.array_equal(d2[:n_in], dl[:n_in]) and np.array_equal(c2[:n_in], co[:n_in])
    for i in range(n_in, n):
        assert np.array_equal(c2[i], co[i % n_in])
        assert np.array_equal(d2[i], (dl[i % n_in].astype(np.float32) * np.float32(n_in) / np.float32(n)).astype(np.float16))


def test_bitfield_mips_are_max_pooled():
    grid = np.zeros(128 ** 3, dtype=np.float32)
    grid[O.lib().orc_morton3D(64, 64, 64)] = 1.0
    bits, mean = O.update_bitfield(grid, 1)
    assert abs(mean - 1.0 / 128 ** 3) < 1e-12
    m = O.lib().orc_morton3D(64, 64, 64)
    assert bits[m // 8] & (1 << (m % 8))
    # mip 1 covers twice the extent: cell (64,64,64)/2 + 32 = (64,64,64)
    m1 = O.lib().orc_morton3D(64, 64, 64)
    assert bits[128 ** 3 // 8 + m1 // 8] & (1 << (m1 % 8))
    assert np.unpackbits(bits[: 128 ** 3 // 8]).sum() == 1


def test_c_abi_exports_every_declared_symbol():
    """The library must load without a GPU and export exactly what include/neus2_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "neus2_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(neus2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    from neus2_b200 import _lib
    lib = _lib.load()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == declared
    # host-only entry points work without a device
    cfg = _lib.Config()
    lib.neus2_default_config(C.byref(cfg))
    assert (cfg.n_levels, cfg.log2_hashmap_size, cfg.base_resolution, cfg.batch_size) == (14, 19, 16, 1 << 18)
    assert abs(cfg.learning_rate - 1e-3) < 1e-9 and abs(cfg.epsilon - 1e-15) < 1e-20 and abs(cfg.ek_loss_weight - 0.01) < 1e-9


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under neus2_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "neus2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/neus_oracle.cpp", ""), f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from neus2_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_lib.Neus2Error):
        _lib.load()


def test_testbed_global_movement_hyperparams():
    """hyperparams of the dynamic-scene schedule as Testbed::reload_network reads them (src/testbed.cu:2118-2139); host logic, no GPU."""
    from neus2_b200.testbed import Testbed, TestbedMode, _rotation_6d_to_matrix
    from oracle import oracle as O
    tb = Testbed(TestbedMode.Nerf)
    tb.reload_network_from_json({"hyperparams": {"predict_global_movement": True}})
    assert tb.predict_global_movement and tb.predict_global_movement_training_step == 300 and tb.finetune_global_movement      # code defaults
    tb.reload_network_from_json({"hyperparams": {"predict_global_movement": True, "predict_global_movement_training_step": 50,
                                                 "finetune_global_movement": False, "next_frame_max_training_step": 1300}})
    assert tb.predict_global_movement_training_step == 50 and not tb.finetune_global_movement and tb.next_frame_max_training_step == 1300
    assert not tb.incremental_reinit_sdf_mlp and tb.incremental_reinit_sdf_mlp_iters == 10
    tb.reload_network_from_json({"hyperparams": {"incremental_reinit_sdf_mlp": True, "incremental_reinit_sdf_mlp_iters": 3}})
    assert tb.incremental_reinit_sdf_mlp and tb.incremental_reinit_sdf_mlp_iters == 3
    tb.reload_network_from_json({"hyperparams": {"predict_global_movement": False, "predict_global_movement_training_step": 50}})
    assert tb.predict_global_movement_training_step == 0                       # :2131-2133: no frozen phase without prediction
    # the host-side 6d -> matrix used by save_transform agrees with the oracle's
    r6 = np.array([1.0, 0.03, -0.02, 0.01, 0.98, 0.05], np.float32).astype(np.float16).astype(np.float32)
    R, _ = O.delta_matrix(r6, [0, 0, 0])
    assert np.abs(_rotation_6d_to_matrix(r6) - R).max() < 2e-7


def test_canonical_training_step_follows_the_reference_counter():
    """m_canonical_training_step is a stateful counter in the reference: reset by training_network_next_frame (src/testbed.cu:2015, 2050)
    and updated at the end of every step (src/testbed_nerf.cu:3510-3527; the "still predicting" branch is a no-op comparison, :3518).
    neus2_canonical_training_step is its closed form; simulate the reference's updates step by step over three frames and compare the
    value each step STARTS with -- the one Testbed::train (grid cadence, testbed.cu:2669), training_prep_nerf (:4011) and
    train_nerf_step (:3785) read."""
    from neus2_b200 import _lib
    f = _lib.load().neus2_canonical_training_step
    for predict, P in ((True, 50), (True, 300), (True, 0), (False, 0)):
        canonical, frame = 0, 0
        for frame, n_steps in enumerate((400, 420, 333)):
            if frame >= 1:
                canonical = 0                                     # training_network_next_frame
            for training_step in range(n_steps):
                assert f(frame, P if predict else 0, training_step) == canonical, (predict, P, frame, training_step)
                after = training_step + 1                         # ++m_training_step (:3510)
                if predict:
                    if frame == 0:
                        canonical = after
                    elif after < P:
                        pass                                      # "==" instead of "=": the counter keeps its value
                    else:
                        canonical = after - P
                else:
                    canonical = after
    # what hangs off it while the canonical model is frozen (frames >= 1, first P steps): the counter stays 0, so the occupancy grid is
    # refreshed on every step (n_prep_to_skip = clamp(0 / 16, 1, 16) = 1) from uniform samples (< 256) and n_rays_total restarts each step
    assert all(f(1, 50, s) == 0 for s in range(51)) and f(1, 50, 51) == 1 and f(1, 50, 306) == 256


# ---- arithmetic the kernels rely on, restated in numpy (fp32, round to nearest even like the device) --------------------------------

def test_chain_positions_from_an_integer_stride_equal_the_serial_additions():
    """sampling.cu: chain_position().  Inside one binade adding the same dt moves the fp32 bit pattern by a constant number of ulps (after the
    first addition, which settles the parity in the tie case), so chain position L of a window is bits(t1) + (L - 1) * stride instead of L
    dependent additions.  Checked against the serial additions for the marching step of the unit cube and for arbitrary steps, over start
    values spread across binades, next to powers of two and next to zero; outside the fast path's condition the kernel adds serially."""
    rs = np.random.RandomState(0)
    ts = np.concatenate([rs.uniform(0, 8, 100000), 2.0 ** rs.randint(-12, 4, 30000) * (1 + rs.uniform(-1e-3, 1e-3, 30000)), rs.uniform(0, 0.01, 30000)]).astype(np.float32)
    ts = ts[ts > 0]
    steps = [np.float32(1.73205080757 / 1024)]
    for _ in range(40):
        b = np.float32(rs.uniform(1e-4, 1e-2)).view(np.uint32)
        steps.append((b & ~np.uint32((1 << rs.randint(0, 12)) - 1)).view(np.float32))      # trailing zero bits: exact ties at some exponents
    for dt in steps:
        seq = np.empty((33, len(ts)), np.float32)
        seq[0] = ts
        for k in range(1, 33):
            seq[k] = seq[k - 1] + dt
        t_far = ts + np.float32(33.0) * dt
        fast = ((t_far.view(np.uint32) >> 23) == (ts.view(np.uint32) >> 23)) & (ts < 256)
        assert fast.mean() > 0.5
        t1 = ts + dt
        t2 = t1 + dt
        stride = t2.view(np.uint32) - t1.view(np.uint32)
        for L in range(1, 33):
            pred = (t1.view(np.uint32) + np.uint32(L - 1) * stride).view(np.float32)
            assert np.array_equal(pred.view(np.uint32)[fast], seq[L].view(np.uint32)[fast]), (float(dt), L)


def test_occupancy_sample_index_is_inverted_exactly():
    """density_grid.cu: k_grid_samples_uniform_ordered visits CELLS and finds the sample that the reference's walk
    idx(i) = ((i + step * n) * 56924617 + 96925573) mod 128^3 (src/testbed_nerf.cu:651) sends there."""
    G3, a, b, a_inv = 128 ** 3, 56924617, 96925573, 2759905401
    assert (a * a_inv) % 2 ** 32 == 1
    rs = np.random.RandomState(1)
    for n in (G3, G3 // 4):
        for step in (0, 1, 7, 12345):
            i = rs.randint(0, n, 20000).astype(np.uint64)
            t = (((i + step * n) * a) % 2 ** 32 + b) % 2 ** 32 % G3                       # the reference's uint32 arithmetic
            back = ((a_inv * ((t - b) % 2 ** 32)) % 2 ** 32 - step * n) % 2 ** 32 % G3
            assert np.array_equal(back, i)
            # and a cell that no sample below n visits maps to an index >= n
            cells = rs.randint(0, G3, 20000).astype(np.uint64)
            pre = ((a_inv * ((cells - b) % 2 ** 32)) % 2 ** 32 - step * n) % 2 ** 32 % G3
            visited = np.isin(cells, (((np.arange(n, dtype=np.uint64) + step * n) * a) % 2 ** 32 + b) % 2 ** 32 % G3) if n < G3 else np.ones(len(cells), bool)
            assert np.array_equal(pre < n, visited)


def test_compositing_with_short_chains_equals_the_serial_loop():
    """loss.cu: composite_chunk().  The reference composites a ray serially (src/testbed_nerf.cu:1567-1627: stop before the first sample whose
    incoming transmittance is below 1e-4; weight = alpha * T; rgb += weight * c; weight_sum += weight; T *= 1 - alpha).  The kernel runs the
    transmittance product over a whole 32-sample chunk first, finds the stop with one ballot (T never grows), forms the weights and products
    per lane and adds them in sample order, with +0 for samples that were not consumed.  Same bits, chunk after chunk."""
    rs = np.random.RandomState(2)
    f32 = np.float32
    for trial in range(300):
        ns = int(rs.randint(1, 200))
        sharp = rs.choice([0.02, 0.2, 0.9])
        alpha = np.clip(rs.beta(0.3, 0.3 / sharp, ns), 0, 1).astype(f32)
        if trial % 7 == 0:
            alpha[rs.randint(0, ns)] = f32(1.0)                        # a fully opaque sample: T drops to exactly 0
        rgb = rs.uniform(0, 1, (ns, 3)).astype(f32)
        # serial reference
        T, ws, acc, kept = f32(1), f32(0), np.zeros(3, f32), 0
        cap_ref = []
        for j in range(ns):
            if T < f32(1e-4):
                break
            w = alpha[j] * T
            acc = acc + w * rgb[j]
            ws = ws + w
            T = T * (f32(1) - alpha[j])
            kept += 1
            cap_ref.append((w, T, acc.copy()))
        # chunked, short chains
        T2, ws2, acc2, kept2, stop = f32(1), f32(0), np.zeros(3, f32), 0, False
        cap = []
        for c0 in range(0, ns, 32):
            if stop:
                break
            cnt = min(32, ns - c0)
            a = np.zeros(32, f32)
            c = np.zeros((32, 3), f32)
            a[:cnt], c[:cnt] = alpha[c0:c0 + cnt], rgb[c0:c0 + cnt]
            keep = f32(1) - a
            T_in, T_out = np.empty(32, f32), np.empty(32, f32)
            t = T2
            for j in range(32):
                T_in[j] = t
                t = t * keep[j]
                T_out[j] = t
            below = np.nonzero((T_in < f32(1e-4)) & (np.arange(32) < cnt))[0]
            consumed = int(below[0]) if len(below) else cnt
            w = a * T_in
            prod = w[:, None] * c
            for j in range(32):
                mine = j < consumed
                acc2 = acc2 + (prod[j] if mine else np.zeros(3, f32))
                ws2 = ws2 + (w[j] if mine else f32(0))
                if mine:
                    cap.append((w[j], T_out[j], acc2.copy()))
            kept2 += consumed
            stop = len(below) > 0
            T2 = T_in[consumed] if stop else t
        assert kept2 == kept
        assert T2.view(np.uint32) == T.view(np.uint32) and ws2.view(np.uint32) == ws.view(np.uint32) and np.array_equal(acc2.view(np.uint32), acc.view(np.uint32))
        for (w1, t1, r1), (w0, t0, r0) in zip(cap, cap_ref):
            assert w1.view(np.uint32) == w0.view(np.uint32) and t1.view(np.uint32) == t0.view(np.uint32) and np.array_equal(r1.view(np.uint32), r0.view(np.uint32))


def test_lookback_prefix_sum_protocol_under_arbitrary_scheduling():
    """lookback.cuh: exclusive prefix sums across the blocks of ONE launch.  A model of the protocol, stepped in random interleavings: tiles are
    numbered by a ticket drawn when the block starts (so every predecessor of a running tile is running or done), a tile publishes
    [epoch | 1 | own sum], walks back over windows of 32 predecessors -- waiting for descriptors of THIS launch's epoch -- until it meets an
    inclusive prefix, then publishes [epoch | 2 | prefix].  The descriptor array is never cleared: it starts out full of the previous
    launches' words (other epochs, both states, arbitrary values), which must not be taken for this launch's."""
    rs = np.random.RandomState(3)

    def run(n_tiles, sums, epoch, desc, max_resident):
        ticket = [0]
        result = {}

        def block():
            tile = ticket[0]
            ticket[0] += 1
            yield                                           # ... the block's own work (pass 1) takes a while
            for _ in range(rs.randint(0, 4)):
                yield
            agg = int(sums[tile])
            if tile == 0:
                desc[0] = (epoch, 2, agg)
                result[0] = 0
                return
            desc[tile] = (epoch, 1, agg)
            excl, p = 0, tile - 1
            while True:
                window = []
                for lane in range(32):
                    idx = p - lane
                    if idx < 0:
                        window.append((2, 0))
                        continue
                    while desc[idx][0] != epoch or desc[idx][1] == 0:
                        yield                               # spin: the predecessor has not published in this launch yet
                    window.append((desc[idx][1], desc[idx][2]))
                    if rs.rand() < 0.3:
                        yield                               # lanes do not read at the same instant
                first = next((k for k, (st, _) in enumerate(window) if st == 2), None)
                last = first if first is not None else 31
                excl += sum(v for _, v in window[:last + 1])
                if first is not None:
                    break
                p -= 32
            desc[tile] = (epoch, 2, excl + agg)
            result[tile] = excl

        pending, running = n_tiles, []
        while pending or running:
            while pending and len(running) < max_resident:   # the hardware starts blocks in any order; tickets are drawn at start
                running.append(block())
                pending -= 1
            g = running[rs.randint(len(running))]
            try:
                next(g)
            except StopIteration:
                running.remove(g)
        return result

    for trial in range(30):
        n_tiles = int(rs.randint(1, 200))
        size = 256
        desc = [(int(rs.randint(0, 5)), int(rs.randint(0, 3)), int(rs.randint(0, 1000))) for _ in range(size)]   # stale words of other launches
        for epoch in (7, 8):                                # two launches in a row on the same, never cleared, array
            sums = rs.randint(0, 50, n_tiles)
            res = run(n_tiles, sums, epoch, desc, max_resident=int(rs.randint(1, 40)))
            want = np.concatenate([[0], np.cumsum(sums)[:-1]])
            assert [res[t] for t in range(n_tiles)] == want.tolist(), (trial, epoch)
            n_tiles = int(rs.randint(1, 200))               # the next launch has another number of tiles
