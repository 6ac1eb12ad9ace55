"""The tcgen05 / tensor-memory facts the training kernel's in-kernel weight gradients rest on, checked on the hardware by tools/umma_probe.cu with
integer data and exact sums: shared-memory descriptor fields and the canonical MN-major no-swizzle layout, the instruction descriptor of
kind::f16, the lane mapping of an M = 64 accumulator, column offsets, N in {16, 32, 48, 64}, accumulation across instructions, the
commit -> mbarrier completion path, and 12 warps x 148 CTAs issuing unordered into one accumulator."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_umma_probe_passes():
    exe = os.path.join(ROOT, "tools", "umma_probe")
    if not os.path.exists(exe):
        pytest.skip("tools/umma_probe not built (python -c 'import __graft_entry__ as g; g.build()')")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASS" in r.stdout and "FAIL" not in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
