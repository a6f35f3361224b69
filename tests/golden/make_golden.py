"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libmaxiref.so, compiled
from /root/reference by oracle/Makefile with -O2 -ffp-contract=off). Run in the build container:

    python tests/golden/make_golden.py

The reference's own tests hold no numeric expectations, so these vectors --
outputs of the reference itself on seeded inputs -- are the pin. Inputs are stored beside the
outputs so the fixtures are self-contained on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from maximilian_b200 import workloads as W      # noqa: E402
from oracle import oracle_py as O               # noqa: E402

KIND = "reference"

CHAINS = [  # (name, osc, filt, env, delay, kwargs)
    ("sinewave_lores", "sinewave", "lores", False, False, {}),
    ("saw_hires", "saw", "hires", False, False, {}),
    ("saw_svf_lp", "saw", "svf", False, False, {}),
    ("phasor_svf_mix", "phasor", "svf", False, False, {"svf_mix": (0.3, 0.2, 0.4, 0.1)}),
    ("saw_biquad_lp", "saw", "biquad", False, False, {}),
    ("triangle_biquad_peak", "triangle", "biquad", False, False, {"biquad_type": "peak"}),
    ("pulse_none", "pulse", "none", False, False, {}),
    ("square_none", "square", "none", False, False, {}),
    ("impulse_none", "impulse", "none", False, False, {}),
    ("coswave_none", "coswave", "none", False, False, {}),
    ("saw_env_delay", "saw", "none", True, True, {"delay_capacity": 96}),
    ("saw_env_lores_delay", "saw", "lores", True, True, {"delay_capacity": 96}),
]


configure = W.configure_bank


def chains():
    V, B, NB = 8, 96, 3
    out = {"V": V, "B": B, "NB": NB}
    for name, osc, filt, env, delay, kw in CHAINS:
        p = W.voice_params(V, seed=1234, delay_size=kw.get("delay_capacity", 96), ragged_delay=True)
        if name == "triangle_biquad_peak":
            p["gain"] = np.linspace(-9.0, 9.0, V)
        b = O.Bank(V, osc=osc, filt=filt, env=env, delay=delay, kind=KIND, **kw)
        configure(b, filt, p, env, delay)
        outs, mixes = [], []
        for blk in range(NB):
            on, off = W.gate(V, B, 4 * blk if blk < 2 else 1)     # two gated blocks, then a silent one
            o, m = b.process(B, on, off, want_mix=True)
            outs.append(o); mixes.append(m)
        out[name + "/out"] = np.stack(outs)
        out[name + "/mix"] = np.stack(mixes)
        out[name + "/phase"] = b.get("phase")
        if delay:
            out[name + "/delay_phase"] = b.get("delay_phase").astype(np.int32)
            out[name + "/ring"] = np.stack([b.ring(v, kw["delay_capacity"]) for v in range(V)])
        if env:
            out[name + "/env_flags"] = b.get("env_flags").astype(np.int32)
            out[name + "/env_amplitude"] = b.get("env_amplitude")
    np.savez_compressed(os.path.join(HERE, "chains.npz"), **out)


MODS = [  # (name, osc, filt, env, delay, which per-sample arrays)
    ("phasorbetween_hires", "phasorbetween", "hires", False, False, ()),
    ("saw_svf_swept_cutoff", "saw", "svf", False, False, ("cutoff",)),
    ("phasor_lores_swept_cutoff_fm", "phasor", "lores", False, False, ("cutoff", "freq")),
    ("saw_env_lores_flanged_delay", "saw", "lores", True, True, ("delay_size",)),
    # per-sample frequency / cutoff on chains that end in a delay line (the modulated instantiations of the delay kernel)
    ("saw_env_delay_fm", "saw", "none", True, True, ("freq",)),
    ("pulse_svf_delay_swept_cutoff_fm", "pulse", "svf", False, True, ("cutoff", "freq")),
]


def mod_arrays(V, B, blk, cap):
    """Per-sample parameter arrays of the modulated cases (deterministic; stored in the fixture as well)."""
    rng = np.random.default_rng(77)
    t = (blk * B + np.arange(B))[:, None] / 48000.0
    centre = 200.0 * np.exp2(4.0 * rng.random(V)); depth = 1.5 * rng.random(V); rate = 0.2 + 6.0 * rng.random(V)
    cutoff = centre[None, :] * np.exp2(depth[None, :] * np.sin(2 * np.pi * rate[None, :] * t))
    carrier = 110.0 * np.exp2(3.0 * rng.random(V)); fdepth = 50.0 * rng.random(V); frate = 0.5 + 8.0 * rng.random(V)
    freq = carrier[None, :] + fdepth[None, :] * np.sin(2 * np.pi * frate[None, :] * t)
    delay = rng.integers(8, cap // 2, V).astype(np.float64); ddepth = rng.random(V); drate = 0.5 + 20.0 * rng.random(V)
    lfo = 2.0 * np.abs(2.0 * ((drate[None, :] * t) % 1.0) - 1.0) - 1.0
    dsize = np.clip(np.floor(delay[None, :] + lfo * ddepth[None, :] * delay[None, :] + 1.0), 1, cap)
    return {"cutoff": cutoff, "freq": freq, "delay_size": dsize}


def mods():
    """SURVEY.md 8(a)-3 phasorBetween and the 8(f) per-sample arguments (frequency, cutoff, delay size), from the reference."""
    V, B, NB, cap = 8, 96, 3, 96
    out = {"V": V, "B": B, "NB": NB, "cap": cap}
    for name, osc, filt, env, delay, which in MODS:
        p = W.voice_params(V, seed=4321, delay_size=cap, ragged_delay=True)
        b = O.Bank(V, osc=osc, filt=filt, env=env, delay=delay, delay_capacity=cap, kind=KIND)
        configure(b, filt, p, env, delay)
        outs, mixes = [], []
        for blk in range(NB):
            on, off = W.gate(V, B, 4 * blk if blk < 2 else 1)
            m = mod_arrays(V, B, blk, cap)
            kw = {k + "_tv": m[k] for k in which}
            for k in which:
                out[f"{name}/{k}_tv/{blk}"] = m[k]
            o, mx = b.process(B, on if env else None, off if env else None, want_mix=True, **kw)
            outs.append(o); mixes.append(mx)
        out[name + "/out"] = np.stack(outs); out[name + "/mix"] = np.stack(mixes)
        out[name + "/phase"] = b.get("phase")
        if delay:
            out[name + "/delay_phase"] = b.get("delay_phase").astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "mods.npz"), **out)


def seeds():
    """The SURVEY.md section 8(c) seed values, regenerated (sr 48000)."""
    out = {}
    b = O.Bank(1, osc="sinewave", filt="lores", kind=KIND); b.set("freq", 440); b.set("cutoff", 1000); b.set("resonance", 2.0)
    out["sine440_lores_1000_2"] = b.process(8)[0].ravel()
    b = O.Bank(1, osc="saw", filt="svf", kind=KIND); b.set("freq", 110); b.set("cutoff", 1000); b.set("resonance", 2.0)
    out["saw110_svf_1000_2"] = b.process(8)[0].ravel()
    b = O.Bank(1, osc="square", delay=True, kind=KIND); b.set("freq", 0); b.set("phase", 0.75)
    b.set("delay_size", 4); b.set("delay_feedback", 0.5)
    out["dl_1_4_05"] = b.process(10)[0].ravel()
    lib = O.load(KIND)
    a, d = lib.mxo_env_attack_coeff(1, 48000), lib.mxo_env_decay_coeff(2, 48000)
    b = O.Bank(1, osc="square", env=True, kind=KIND); b.set("freq", 0); b.set("phase", 0.75)
    b.set("env_attack", a); b.set("env_decay", d); b.set("env_sustain", .5); b.set("env_release", d); b.set("env_holdtime", 1)
    out["adsr_1_2_05_2"] = b.process(14, [0], [6])[0].ravel()
    out["adsr_coeffs"] = np.array([a, d])
    np.savez_compressed(os.path.join(HERE, "seeds.npz"), **out)


def spectral():
    C, n, hop = 2, 1024, 512
    x = W.channel_streams(C, 5 * hop, seed=4242)
    st = O.Stft(C, n, hop, kind=KIND)
    r = st.process(x)
    mf = O.Mfcc(512, 42, 40, 20.0, 20000.0, 48000, kind=KIND)
    co, mb = mf.process(r["mags"])
    mf13 = O.Mfcc(512, 42, 13, 20.0, 20000.0, 44100, kind=KIND)
    co13, _ = mf13.process(r["mags"])
    y = O.Istft(C, n, hop, kind=KIND).process(r["mags"], r["phases"])
    # hop 256 variant, short
    x2 = W.channel_streams(1, 1024 + 3 * 256, seed=99)
    r2 = O.Stft(1, 1024, 256, kind=KIND).process(x2)
    np.savez_compressed(os.path.join(HERE, "spectral.npz"), x=x, window=st.window(), re=r["re"], im=r["im"],
                        mags=r["mags"], phases=r["phases"], mfcc40=co, melbands=mb, mfcc13_44k=co13, istft=y,
                        x_hop256=x2, mags_hop256=r2["mags"], re_hop256=r2["re"], im_hop256=r2["im"])


def tables():
    """sineBuffer[514], transition[1001] and the double in front of sineBuffer (read by sinebuf4 on its wrap sample) as the
    compiled reference holds them: data the product and the C port are handed at run time (mxb_ctx_set_tables / mxo_set_tables)."""
    s, t, b = O.get_tables(KIND)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), sine=s, transition=t, sine_before=np.float64(b))


def patches():
    """Voice patches (tests/patch_cases.py) run by the reference's own objects: outputs, buses and inputs of 3 blocks."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import patch_cases as PC
    V, B, NB = 12, 160, 3
    out = {"V": V, "B": B, "NB": NB}
    for name, d, params, inputs, exact, taps in PC.cases():
        p = O.Patch(d, V, delay_taps=taps, kind=KIND)
        for k, v in params(V, 2468).items():
            p.set(k, v)
        outs, mixes = [], []
        for blk in range(NB):
            ins = inputs(V, B, blk, 1357)
            for k, a in ins.items():
                out[f"{name}/in/{k}/{blk}"] = a
            o, m = p.process(B, ins, want_mix=True)
            outs.append(o); mixes.append(m)
        out[name + "/out"] = np.stack(outs); out[name + "/mix"] = np.stack(mixes)
    np.savez_compressed(os.path.join(HERE, "patches.npz"), **out)


def chorus():
    """maxiChorus run by the reference's own objects, drawing its noise from libc rand() after srand(seed); the fixture keeps the noise
    stream (what maxiOsc::noise() returned, in frame-major / voice-minor order) next to the outputs, so that the replay needs no rand()."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import patch_cases as PC
    V, B, NB, seed = 10, 200, 3, 4242
    name, d, params, _, exact, taps = PC.chorus()
    noise = O.noise_fill(seed, NB * B * V, KIND).reshape(NB, B, V)
    p = O.Patch(d, V, delay_taps=taps, kind=KIND)
    for k, v in params(V, 99).items():
        p.set(k, v)
    O.srand(seed, KIND)
    outs, mixes = [], []
    for blk in range(NB):
        o, m = p.process(B, {"noise": noise[blk]}, want_mix=True)       # the reference ignores the stream: it draws the same values itself
        outs.append(o); mixes.append(m)
    np.savez_compressed(os.path.join(HERE, "chorus.npz"), V=V, B=B, NB=NB, taps=taps, noise=noise, out=np.stack(outs), mix=np.stack(mixes))


if __name__ == "__main__":
    O.build("reference")
    if len(sys.argv) > 1 and sys.argv[1] == "chorus":      # add this fixture without rewriting the others
        chorus()
    elif len(sys.argv) > 1 and sys.argv[1] == "mods":
        mods()
    else:
        chains(); seeds(); spectral(); mods(); tables(); patches(); chorus()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
