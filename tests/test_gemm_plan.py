"""Host logic of the tcgen05 GEMM tile plan (fo_debug_plan needs no GPU): invariants the kernels rely on, and the plans of
the shapes the measured numbers in DESIGN.md were taken with."""
import ctypes as C

import pytest

from freeze_omni_b200 import _lib

SMS = 148


def plan(rows, n_out, k, defer=0):
    lib = _lib.load()
    s, b, p = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.fo_debug_plan(rows, n_out, k, defer, C.byref(s), C.byref(b), C.byref(p)))
    return s.value, b.value, p.value


@pytest.mark.parametrize("rows", [4, 28, 64, 256, 384, 512, 1024])
@pytest.mark.parametrize("n_out,k", [(3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096), (1024, 19456), (3584, 2048), (2048, 5120)])
@pytest.mark.parametrize("defer", [0, 1])
def test_skinny_plan_invariants(rows, n_out, k, defer):
    swap, bn, split = plan(rows, n_out, k, defer)
    kblocks = k // 64
    assert swap == 1                                             # weights on the 128-row UMMA-M side
    assert bn % 16 == 0 and 16 <= bn <= 256                      # a legal UMMA N
    assert 1 <= split <= 8 and split <= kblocks
    ta, tb = -(-n_out // 128), -(-rows // bn)
    if bn < 256 and tb > 1:
        assert ta * tb * min(split, 4) <= 2 * SMS                # at most two CTAs per SM in flight
    if not defer:
        assert split <= 4 or ta * tb * split <= SMS              # an in-GEMM reduction only grows past 4 to fill one wave


def test_plans_of_the_measured_step():
    # 64 sessions (256 rows): QKV / FFN1 un-split at 32-token slices; out-proj and FFN2 (reduction deferred to the LayerNorm)
    assert plan(256, 3072, 1024) == (1, 32, 1)
    assert plan(256, 4096, 1024) == (1, 32, 1)
    assert plan(256, 1024, 1024, 1) == (1, 64, 4)                # short K: one fat CTA per SM (r02 in-chain sweep)
    assert plan(256, 1024, 4096, 1) == (1, 32, 4)
    # one session: K of the QKV GEMM split for the attention kernel to sum
    s, b, p = plan(4, 3072, 1024, 1)
    assert (s, b) == (1, 16) and p >= 2


def test_fat_plan():
    swap, bn, split = plan(23936, 4096, 1024)
    assert split == 1 and 64 <= bn <= 256 and bn % 16 == 0
    swap, bn, split = plan(5120, 1024, 9216)                     # conv2 of a 64-session step: wave-quantised UMMA N
    assert split == 1 and bn % 16 == 0
