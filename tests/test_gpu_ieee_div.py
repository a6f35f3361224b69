"""The straight-line IEEE division / reciprocal the modulated kernels use for maxiOsc's per-sample increment 1./(sampleRate/frequency)
(csrc/bank_kernels.cuh: div_rn_unchecked, rcp_rn_unchecked, osc_increment) against the compiler's own `/` operators on the same device,
bit for bit -- and against the host's IEEE division, which is what the reference computes."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "cuda", "libmxbselftest.so")


def test_selftest_library_is_built():
    """__graft_entry__.build() compiles it next to the product library (it travels to the GPU box like that one)."""
    from maximilian_b200 import build
    assert os.path.exists(build.build_selftest())


def _run(a, b):
    lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    lib.mxbtest_div.restype = C.c_int
    lib.mxbtest_div.argtypes = [dp, dp, C.c_longlong] + [dp] * 6
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    outs = [np.empty_like(a) for _ in range(6)]
    rc = lib.mxbtest_div(a.ctypes.data_as(dp), b.ctypes.data_as(dp), a.size, *[o.ctypes.data_as(dp) for o in outs])
    assert rc == 0, f"CUDA error {rc}"
    return outs


def _same_bits(x, y):
    return np.array_equal(x.view(np.uint64), y.view(np.uint64))


@pytest.mark.gpu
def test_unchecked_sequences_are_the_operators_on_sane_operands():
    rng = np.random.default_rng(77)
    n = 1 << 24
    # numerators like sample rates, denominators like frequencies, plus the whole range freq_sane() admits (log-uniform, both signs)
    a = np.concatenate([np.full(n // 4, 48000.0), np.full(n // 4, 44100.0), rng.uniform(8000.0, 768000.0, n // 4),
                        2.0 ** rng.uniform(-20, 40, n // 4)])
    b = np.concatenate([27.5 * 2.0 ** rng.uniform(0, 9.5, n // 4), rng.uniform(0.01, 24000.0, n // 4), 2.0 ** rng.uniform(-60, 60, n // 2)])
    b *= rng.choice([-1.0, 1.0], n, p=[0.1, 0.9])
    q_fn, q_op, r_fn, r_op, i_fn, i_op = _run(a, b)
    assert _same_bits(q_fn, q_op) and _same_bits(r_fn, r_op) and _same_bits(i_fn, i_op)
    with np.errstate(all="ignore"):
        assert _same_bits(q_op, a / b) and _same_bits(r_op, 1.0 / b) and _same_bits(i_op, 1.0 / (a / b))      # IEEE on both sides


@pytest.mark.gpu
def test_increment_falls_back_to_the_operators_outside_the_sane_range():
    """Zero, denormal, huge, infinite and NaN frequencies (and sample rates): osc_increment() must answer what 1./(sr/f) answers."""
    sp = np.array([0.0, -0.0, 5e-324, 1e-310, 2.0 ** -1022, 2.0 ** -61, 2.0 ** 61, 1e300, 1.7e308, np.inf, -np.inf, np.nan, 440.0, -440.0, 1e-30, 1e30])
    a, b = np.meshgrid(np.array([48000.0, 0.0, 1e-8, 2.0 ** 41, np.inf, np.nan, 1.0]), sp, indexing="ij")
    _, _, _, _, i_fn, i_op = _run(a.ravel(), b.ravel())
    assert np.array_equal(np.isnan(i_fn), np.isnan(i_op))
    m = ~np.isnan(i_op)
    assert _same_bits(i_fn[m], i_op[m])
