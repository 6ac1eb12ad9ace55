"""CPU pin of the batch-size controller: neus2_next_rays_per_batch (C ABI, host arithmetic used by the training step) against the REFERENCE's own
Testbed::Nerf::Training::Counters::update_after_training (src/testbed_nerf.cu:3412-3438 with the struct of testbed.h:599-616, streamed into a host
build by oracle/Makefile.ref -> oracle/_ref/libref_counters_host.so).  The table below was produced by that code; where the reference tree is
mounted it is recomputed, together with a few thousand random cases."""
import ctypes as C
import os

import numpy as np
import pytest

from neus2_b200 import _lib

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_counters_host.so")
HAVE_REF = os.path.exists(LIB) and os.path.isdir("/root/reference/src")

# (rays_per_batch, target_batch_size, numsteps_counter, numsteps_counter_compacted) -> (rays_per_batch next, measured, measured before compaction)
TABLE = [
    ((4096, 262144, 631000, 100000), (10752, 100000, 631000)),
    ((4096, 262144, 40000, 16000), (67200, 16000, 40000)),
    ((16384, 262144, 700000, 262144), (16384, 262144, 700000)),
    ((16384, 262144, 700000, 262145), (16384, 262145, 700000)),
    ((16385, 262144, 700000, 262144), (16512, 262144, 700000)),
    ((200000, 262144, 5000, 1000), (262144, 1000, 5000)),
    ((4096, 262144, 0, 0), (4096, 0, 0)),
    ((4096, 262144, 100, 0), (4096, 0, 0)),
    ((12345, 16384, 50000, 33333), (6144, 33333, 50000)),
    ((262144, 262144, 900000, 300000), (229120, 300000, 900000)),
    ((128, 262144, 3, 1), (262144, 1, 3)),
]


def _reference(rpb, target, n, nc, losses=None):
    L = C.CDLL(LIB)
    out, sc = (C.c_uint32 * 4)(), (C.c_float * 3)()
    ptr = [None, None, None] if losses is None else [a.ctypes.data_as(C.c_void_p) for a in losses]
    L.refc_update_after_training(C.c_uint32(rpb), C.c_uint32(target), C.c_uint32(n), C.c_uint32(nc), ptr[0], ptr[1], ptr[2], C.c_int(0 if losses is None else 1), out, sc)
    return (int(out[0]), int(out[1]), int(out[2])), (float(sc[0]), float(sc[1]), float(sc[2]))


def _mine(rpb, target, n, nc):
    if n == 0 or nc == 0:                # update_after_training returns before touching rays_per_batch (:3420-3422); so does the step (csrc/api.cu)
        return rpb, 0, 0
    return int(_lib.load().neus2_next_rays_per_batch(rpb, target, nc, 1, 1 << 18)), nc, n


def test_next_rays_per_batch_matches_the_reference_table():
    for args, want in TABLE:
        assert _mine(*args) == want, args
    L = _lib.load()
    # data-parallel contexts round to 128 rays PER RANK so that every rank traces the same number of rays
    assert L.neus2_next_rays_per_batch(4096, 262144, 100000, 8, 1 << 18) == 11264 and L.neus2_next_rays_per_batch(4096, 262144, 100000, 1, 8192) == 8192


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference tree (oracle/_ref/libref_counters_host.so)")
def test_table_and_random_cases_against_the_reference_code():
    for args, want in TABLE:
        assert _reference(*args)[0] == want, args
    rs = np.random.RandomState(0)
    for _ in range(3000):
        rpb = int(rs.randint(1, (1 << 18) + 1)); target = int(rs.choice([1 << 14, 1 << 16, 1 << 18, 250000]))
        nc = int(rs.randint(0, 400000)); n = nc + int(rs.randint(0, 600000))
        assert _reference(rpb, target, n, nc)[0] == _mine(rpb, target, n, nc), (rpb, target, n, nc)
    # the reported loss scalars: sum over the rays * measured / target, left to right (:3429-3431; csrc/api.cu forms the same expression on its own sums)
    loss = [rs.uniform(0, 1e-3, 4096).astype(np.float32) for _ in range(3)]
    (_, m, _), sc = _reference(4096, 262144, 631000, 100000, loss)
    for k in range(3):
        s = np.float32(0)
        for v in loss[k]:
            s = np.float32(s + v)
        assert np.float32(sc[k]) == np.float32(np.float32(s * np.float32(m)) / np.float32(262144))
