"""Shared golden-vector checks: the same fixtures (tests/golden/*.npz, generated from the compiled
reference by tests/golden/make_golden.py) are replayed through the plain-C oracle on the CPU
(test_golden.py, exact) and through the CUDA path on the GPU (test_gpu_*.py, tolerances below).

Tolerances:
  fp64 sample values   |g - r| <= 1e-5*|r| + 1e-12   (north_star's 1e-5 relative; expected <= 1e-12)
  integer state        exact
  FFT re/im/magnitude  bit-equal expected; acceptance |d| <= 1e-5 * max|mag| per frame
"""
import importlib.util
import os

import numpy as np

from maximilian_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load_make_golden():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def chain_cases():
    return _load_make_golden().CHAINS


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def assert_samples_close(got, ref, exact=False, what=""):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if exact:
        assert np.array_equal(got, ref, equal_nan=True), what
        return 0.0
    assert np.array_equal(np.isnan(got), np.isnan(ref)), what
    err = np.abs(got - ref)
    tol = 1e-5 * np.abs(ref) + 1e-12
    bad = np.nan_to_num(err - tol, nan=-1.0) > 0
    assert not bad.any(), (what, float(np.nanmax(err)), int(bad.sum()))
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.nanmax(np.where(np.abs(ref) > 1e-9, err / np.abs(ref), 0.0)) if err.size else 0.0
    return float(rel)


def run_chain_case(make_bank, case, exact, exact_mix=None):
    """make_bank(V, osc=, filt=, env=, delay=, **kw) -> object with set/get/process/ring like oracle_py.Bank."""
    name, osc, filt, env, delay, kw = case
    exact_mix = exact if exact_mix is None else exact_mix
    g = load("chains")
    V, B, NB = int(g["V"]), int(g["B"]), int(g["NB"])
    p = W.voice_params(V, seed=1234, delay_size=kw.get("delay_capacity", 96), ragged_delay=True)
    if name == "triangle_biquad_peak":
        p["gain"] = np.linspace(-9.0, 9.0, V)
    b = make_bank(V, osc=osc, filt=filt, env=env, delay=delay, **kw)
    W.configure_bank(b, filt, p, env, delay)
    worst = 0.0
    for blk in range(NB):
        on, off = W.gate(V, B, 4 * blk if blk < 2 else 1)
        o, m = b.process(B, on, off, want_mix=True)
        worst = max(worst, assert_samples_close(o, g[name + "/out"][blk], exact, f"{name} out blk{blk}"))
        # the mix is a sum over voices: order of summation is free on the GPU (fp64 reassociation)
        if exact_mix:
            assert np.array_equal(m, g[name + "/mix"][blk], equal_nan=True)
        else:
            np.testing.assert_allclose(m, g[name + "/mix"][blk], rtol=1e-9, atol=1e-12)
    assert_samples_close(b.get("phase"), g[name + "/phase"], exact, name + " phase")
    if delay:
        assert np.array_equal(b.get("delay_phase").astype(np.int32), g[name + "/delay_phase"]), name   # integer: exact
        ring = np.stack([b.ring(v, kw["delay_capacity"]) for v in range(V)])
        assert_samples_close(ring, g[name + "/ring"], exact, name + " ring")
    if env:
        assert np.array_equal(b.get("env_flags").astype(np.int32), g[name + "/env_flags"]), name
        assert_samples_close(b.get("env_amplitude"), g[name + "/env_amplitude"], exact, name + " amp")
    return worst


def mod_cases():
    return _load_make_golden().MODS


def run_mod_case(make_bank, case, exact, rtol=1e-9):
    """Replays one tests/golden/mods.npz case (phasorBetween / per-sample frequency, cutoff, delay size). `exact`:
    compare bits; else |g - r| <= rtol*|r| + 1e-12 (device libm in the per-sample coefficient design)."""
    name, osc, filt, env, delay, which = case
    g = load("mods")
    V, B, NB, cap = int(g["V"]), int(g["B"]), int(g["NB"]), int(g["cap"])
    p = W.voice_params(V, seed=4321, delay_size=cap, ragged_delay=True)
    b = make_bank(V, osc=osc, filt=filt, env=env, delay=delay, delay_capacity=cap)
    W.configure_bank(b, filt, p, env, delay)
    for blk in range(NB):
        on, off = W.gate(V, B, 4 * blk if blk < 2 else 1)
        kw = {k + "_tv": g[f"{name}/{k}_tv/{blk}"] for k in which}
        o, m = b.process(B, on if env else None, off if env else None, want_mix=True, **kw)
        if exact:
            assert np.array_equal(o, g[name + "/out"][blk], equal_nan=True), f"{name} blk{blk}"
        else:
            np.testing.assert_allclose(o, g[name + "/out"][blk], rtol=rtol, atol=1e-12, err_msg=f"{name} blk{blk}")
        np.testing.assert_allclose(m, g[name + "/mix"][blk], rtol=max(rtol, 1e-9), atol=1e-11)
    if exact:
        assert np.array_equal(b.get("phase"), g[name + "/phase"]), name
    if delay:
        assert np.array_equal(b.get("delay_phase").astype(np.int32), g[name + "/delay_phase"]), name
