"""Host logic of the one-launch layer stack (csrc/fo_stack.cu; fo_debug_stack_plan needs no GPU): which streaming steps
qualify and the shared-memory plan the kernel runs with on a B200 (148 SMs, 227 KB of opt-in shared memory per CTA)."""
import ctypes as C

import pytest

from freeze_omni_b200 import _lib

SMS, SMEM = 148, 232448


def plan(d, ff, h, n, t, window=64, layers=24, sms=SMS, smem=SMEM):
    lib = _lib.load()
    out = [C.c_int() for _ in range(5)]
    rc = lib.fo_debug_stack_plan(d, ff, h, n, t, window, layers, sms, smem, *[C.byref(o) for o in out])
    return rc, [o.value for o in out]


def test_shipped_one_session_plan():
    rc, (smem, kc, rq, rf, ro) = plan(1024, 4096, 16, 1, 4)
    assert rc == 0
    assert (rq, rf, ro) == (21, 28, 7)               # weight rows per CTA: 3D / FF / D over 148 SMs, rounded up
    assert kc == 4096                                # 4 rows x 4096 fp16 of FFN2 activations fit the activation buffer at once
    assert 180 * 1024 < smem <= SMEM                 # three weight slots + activations + partial tiles + attention tile: one CTA per SM


@pytest.mark.parametrize("n,t,kc", [(2, 4, 2048), (4, 4, 1024), (2, 7, 1024)])
def test_ffn2_activations_are_chunked_beyond_four_rows(n, t, kc):
    rc, (smem, chunk, _, _, _) = plan(1024, 4096, 16, n, t)
    assert rc == 0 and chunk == kc and chunk % 128 == 0 and 4096 % chunk == 0
    assert smem <= SMEM


def test_steps_that_do_not_qualify():
    assert plan(1024, 4096, 16, 5, 4)[0] == 1        # 20 token rows > 16
    assert plan(1024, 4096, 16, 1, 9)[0] == 1        # more than 8 frames per call
    assert plan(1024, 4096, 8, 1, 4)[0] == 1         # heads x 64 != d_model
    assert plan(1536, 4096, 24, 1, 4)[0] == 1        # d_model > 1024
    assert plan(1024, 4096, 16, 1, 4, window=128)[0] == 1      # window + frames > 128 keys
    assert plan(1024, 4096, 16, 1, 4, sms=8)[0] == 1           # 16 attention units do not fit 8 CTAs
    assert plan(1024, 4096, 16, 1, 4, smem=100 * 1024)[0] == 1  # not enough shared memory


def test_toy_widths_qualify():
    rc, (smem, kc, rq, rf, ro) = plan(128, 256, 2, 4, 4, layers=2)
    assert rc == 0 and kc == 256 and (rq, rf, ro) == (3, 2, 1)
