"""Voice patches used by the parity tests: each case is (name, PatchDef, parameter dict builder, input builder, exact, taps).

`exact`: every stage of the patch is arithmetic both sides round identically (bit parity expected on the GPU); else the
patch contains a stage that designs coefficients or calls pow/atan/tan/cos per sample (libdevice vs glibc): 1e-9 relative."""
import numpy as np

from maximilian_b200.patchdef import PatchDef, R, ENVGEN_HOLD
from maximilian_b200 import workloads as W


def _trig_stream(V, B, blk, seed, density=0.004, hold=(5, 300)):
    """per-sample trigger [B][V]: note gates that open and close anywhere inside a block, several per block for some voices"""
    rng = np.random.default_rng(seed + 31 * blk)
    t = np.zeros((B, V))
    for v in range(V):
        pos = 0
        while True:
            pos += int(rng.geometric(density))
            if pos >= B:
                break
            ln = int(rng.integers(hold[0], hold[1]))
            t[pos:pos + ln, v] = 1.0
            pos += ln + 1
    return t


def polysynth():
    """One voice of cpp/commandline/maximilian_examples/15.polysynth/main.cpp:54-70: two pulse VCOs (the second detuned by a
    sinebuf LFO) summed into a lores VCF whose cutoff follows pitch + LFO, multiplied by the ADSR AFTER the filter."""
    d = W.polysynth_patch()

    def params(V, seed):
        return W.polysynth_params(V, seed)

    def inputs(V, B, blk, seed):
        return dict(trigger=_trig_stream(V, B, blk, seed))
    return ("polysynth", d, params, inputs, False, 0)


def family_exact():
    """Table oscillators, one-pole filters, DC blocker and the two rational nonlinearities: every stage rounds the same way
    on both sides (floor / table look-ups / + - * / only)."""
    d = PatchDef()
    d.stage("osc", d.P("freq"), kind="sinebuf4", dst=R(0))
    d.stage("osc", d.P("freq2"), kind="sawn", dst=R(1))
    d.stage("osc", d.P("freq"), kind="sinebuf", dst=R(2))
    d.stage("filter", R(0), d.P("c01"), kind="lopass", dst=R(3))
    d.stage("filter", R(1), d.P("c01"), kind="hipass", dst=R(4))
    d.stage("add", R(3), R(4), dst=R(5))
    d.stage("add", R(5), R(2), dst=R(5))
    d.stage("dcblock", R(5), d.K(0.995), dst=R(6))
    d.stage("nonlin", R(6), kind="fastatan", dst=R(7))
    d.stage("nonlin", R(7), d.P("shape"), kind="fastatandist", dst=R(8))
    d.stage("nonlin", R(8), kind="hardclip", dst=R(9))
    d.stage("out", R(9))
    d.stage("mix_stereo", R(9), d.P("pan"))

    def params(V, seed):
        p = W.voice_params(V, seed=seed)
        rng = np.random.default_rng(seed)
        return dict(freq=p["freq"], freq2=p["freq"] * 0.5 + 20.0, c01=rng.random(V), shape=1.0 + 9.0 * rng.random(V), pan=p["pan"])
    return ("family_exact", d, params, lambda V, B, blk, seed: {}, True, 0)


def family_libm():
    """bandpass, maxiSVF and maxiBiquad designed on every sample from modulated arguments, atan / pow nonlinearities."""
    d = PatchDef()
    d.stage("osc", d.K(3.0), kind="sinewave", dst=R(0))                                   # LFO
    d.stage("osc", d.P("freq"), kind="saw", dst=R(1))
    d.stage("mul", R(0), d.K(400.0), dst=R(2))
    d.stage("add", d.P("cutoff"), R(2), dst=R(2))                                         # cutoff + 400*LFO
    d.stage("filter", R(1), R(2), d.P("res01"), kind="bandpass", dst=R(3))
    d.stage("svf", R(1), R(2), d.P("res"), d.K(0.3), d.K(0.2), d.K(0.4), d.K(0.1), dst=R(4))
    d.stage("biquad", R(1), R(2), d.P("q"), d.K(3.0), kind="peak", dst=R(5))
    d.stage("add", R(3), R(4), dst=R(6))
    d.stage("add", R(6), R(5), dst=R(6))
    d.stage("mul", R(6), d.K(0.15), dst=R(6))                                             # keep the distortions off their clip rails
    d.stage("nonlin", R(6), d.P("shape"), kind="atandist", dst=R(7))
    d.stage("nonlin", R(7), kind="softclip", dst=R(8))
    d.stage("nonlin", R(8), d.K(1.5), d.K(0.7), kind="asymclip", dst=R(9))
    d.stage("out", R(9))

    def params(V, seed):
        p = W.voice_params(V, seed=seed)
        rng = np.random.default_rng(seed)
        return dict(freq=p["freq"], cutoff=600.0 + p["cutoff"], res01=0.1 + 0.85 * rng.random(V), res=p["res_svf"], q=p["q_biquad"],
                    shape=1.0 + 5.0 * rng.random(V))
    return ("family_libm", d, params, lambda V, B, blk, seed: {}, False, 0)


def envgen_flanger():
    """maxiEnvGen (attack / decay / HOLD / release segments with curves, per-sample trigger) shaping a triangle oscillator,
    through maxiFlanger (LFO-swept delay size), maxiDelayline::dl and dlFromPosition; maxiEnv::ar beside it."""
    d = PatchDef()
    d.envgen([0.0, 1.0, 0.6, 0.6, 0.0], [12.0, 30.0, ENVGEN_HOLD, 45.0], [1.0, 2.0, 1.0, 0.5], loop=False, retrigger=True)
    d.stage("envgen", d.IN("gate"), dst=R(0))
    d.stage("osc", d.P("freq"), kind="triangle", dst=R(1))
    d.stage("mul", R(1), R(0), dst=R(2))
    d.stage("flanger", R(2), d.K(180.0), d.P("fb"), d.K(0.7), d.K(0.6), dst=R(3))
    d.stage("delay", R(3), d.P("size"), d.P("fb"), kind="dl", dst=R(4))
    d.stage("delay", R(2), d.P("size"), d.P("fb"), d.P("pos"), kind="position", dst=R(5))
    d.stage("env_ar", R(1), d.IN("trig01"), d.P("att"), d.P("rel"), d.K(40.0), dst=R(6))
    d.stage("add", R(4), R(5), dst=R(7))
    d.stage("add", R(7), R(6), dst=R(7))
    d.stage("out", R(7))
    d.stage("mix_stereo", R(7), d.P("pan"))

    def params(V, seed):
        p = W.voice_params(V, seed=seed)
        rng = np.random.default_rng(seed)
        att, dec, rel = W.env_coeffs(p)
        return dict(freq=p["freq"], fb=p["delay_feedback"], size=rng.integers(2, 300, V).astype(np.float64),
                    pos=rng.integers(0, 320, V).astype(np.float64), att=att, rel=rel, pan=p["pan"])

    def inputs(V, B, blk, seed):
        g = _trig_stream(V, B, blk, seed, density=0.003, hold=(20, 400))
        return dict(gate=2.0 * g - 1.0, trig01=_trig_stream(V, B, blk, seed + 5))       # the gate swings -1 / +1: maxiEnvGen watches zero crossings
    return ("envgen_flanger", d, params, inputs, False, 512)        # pow(level, curve) per sample: libdevice vs glibc


def chorus(modulated_speed=False):
    """maxiChorus (src/maximilian.h:1180-1212) on a saw: lores-filtered noise sweeps two delay lines. The noise is an input stream -- what
    maxiOsc::noise() returns, sample by sample, in (frame, voice) order; the compiled reference draws it itself from libc rand()
    (oracle_py.noise_fill / srand keep the two in step). modulated_speed: the chorus speed follows a slow LFO (the lores design then
    runs per sample instead of once per block)."""
    d = PatchDef()
    d.stage("osc", d.P("freq"), kind="saw", dst=R(0))
    d.stage("mul", R(0), d.K(0.5), dst=R(0))
    if modulated_speed:
        d.stage("osc", d.K(2.0), kind="triangle", dst=R(2))
        d.stage("mul", R(2), d.K(15.0), dst=R(2))
        d.stage("add", R(2), d.K(40.0), dst=R(2))                                     # 25 .. 55 Hz
        speed = R(2)
    else:
        speed = d.K(35.0)
    d.stage("chorus", R(0), d.K(120.0), d.P("fb"), speed, d.P("depth"), d.IN("noise"), dst=R(1))
    d.stage("out", R(1))
    d.stage("mix_stereo", R(1), d.P("pan"))

    def params(V, seed):
        p = W.voice_params(V, seed=seed)
        rng = np.random.default_rng(seed + 1)
        return dict(freq=p["freq"], fb=p["delay_feedback"], depth=0.2 + 0.6 * rng.random(V), pan=p["pan"])
    return ("chorus_lfo" if modulated_speed else "chorus", d, params, None, False, 512)        # inputs: the noise stream, made by the test


def cases():
    return [polysynth(), family_exact(), family_libm(), envgen_flanger()]
