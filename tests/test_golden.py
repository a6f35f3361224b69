"""The plain-C oracle against the committed golden vectors (CPU, exact).

The fixtures were produced by the compiled, unmodified reference (tests/golden/make_golden.py);
this test runs wherever gcc does -- including the GPU box, where /root/reference is absent --
and is what keeps the oracle pinned there.
"""
import numpy as np
import pytest

import golden_checks as G


@pytest.mark.parametrize("case", G.chain_cases(), ids=lambda c: c[0])
def test_chain_golden_exact(port, case):
    G.run_chain_case(lambda V, **kw: port.Bank(V, kind="port", **kw), case, exact=True)


@pytest.mark.parametrize("case", G.mod_cases(), ids=lambda c: c[0])
def test_modulated_golden_exact(port, case):
    """phasorBetween and the per-sample frequency / cutoff / delay-size arguments (tests/golden/mods.npz)."""
    G.run_mod_case(lambda V, **kw: port.Bank(V, kind="port", **kw), case, exact=True)


def test_survey_seed_values(port):
    g = G.load("seeds")
    # literal values quoted in SURVEY.md section 8(c)
    assert g["sine440_lores_1000_2"][1] == 0.00098493646104162898
    assert g["saw110_svf_1000_2"][3] == 0.00034993828994576504
    np.testing.assert_array_equal(g["dl_1_4_05"], [0, 0, 0, 0, .25, .25, .25, .25, .375, .375])
    assert g["adsr_coeffs"][0] == 0.091482424348313218 and g["adsr_coeffs"][1] == 0.95316188323478757

    b = port.Bank(1, osc="sinewave", filt="lores"); b.set("freq", 440); b.set("cutoff", 1000); b.set("resonance", 2.0)
    np.testing.assert_array_equal(b.process(8)[0].ravel(), g["sine440_lores_1000_2"])
    b = port.Bank(1, osc="saw", filt="svf"); b.set("freq", 110); b.set("cutoff", 1000); b.set("resonance", 2.0)
    np.testing.assert_array_equal(b.process(8)[0].ravel(), g["saw110_svf_1000_2"])
    b = port.Bank(1, osc="square", delay=True); b.set("freq", 0); b.set("phase", 0.75)
    b.set("delay_size", 4); b.set("delay_feedback", 0.5)
    np.testing.assert_array_equal(b.process(10)[0].ravel(), g["dl_1_4_05"])
    lib = port.load("port")
    a, d = lib.mxo_env_attack_coeff(1, 48000), lib.mxo_env_decay_coeff(2, 48000)
    assert (a, d) == tuple(g["adsr_coeffs"])
    b = port.Bank(1, osc="square", env=True); b.set("freq", 0); b.set("phase", 0.75)
    b.set("env_attack", a); b.set("env_decay", d); b.set("env_sustain", .5); b.set("env_release", d); b.set("env_holdtime", 1)
    np.testing.assert_array_equal(b.process(14, [0], [6])[0].ravel(), g["adsr_1_2_05_2"])


def test_spectral_golden_exact(port):
    g = G.load("spectral")
    st = port.Stft(2, 1024, 512)
    np.testing.assert_array_equal(st.window(), g["window"])
    r = st.process(g["x"])
    for k in ("re", "im", "mags", "phases"):
        np.testing.assert_array_equal(r[k], g[k])
    co, mb = port.Mfcc(512, 42, 40, 20.0, 20000.0, 48000).process(g["mags"])
    np.testing.assert_array_equal(co, g["mfcc40"])
    np.testing.assert_array_equal(mb, g["melbands"])
    co13, _ = port.Mfcc(512, 42, 13, 20.0, 20000.0, 44100).process(g["mags"])
    np.testing.assert_array_equal(co13, g["mfcc13_44k"])
    y = port.Istft(2, 1024, 512).process(g["mags"], g["phases"])
    np.testing.assert_array_equal(y, g["istft"])
    r2 = port.Stft(1, 1024, 256).process(g["x_hop256"])
    np.testing.assert_array_equal(r2["mags"], g["mags_hop256"])
    np.testing.assert_array_equal(r2["re"], g["re_hop256"])


def test_reference_fft_quirks(port):
    """Properties of the reference transform the CUDA path has to reproduce."""
    g = G.load("spectral")
    x = g["x"][0, :1024].astype(np.float64) * 0  # frame 0 = 512 zeros + first 512 samples
    frame = np.concatenate([np.zeros(512), g["x"][0, :512].astype(np.float64)]) * g["window"].astype(np.float64)
    X = np.fft.rfft(frame)
    re, im = g["re"][0, 0].astype(np.float64), g["im"][0, 0].astype(np.float64)
    scale = np.abs(X).max()
    # forward transform = complex conjugate of the textbook DFT, bins 1..255 and 257..511
    k = np.r_[1:256, 257:512]
    assert np.abs(re[k] - X.real[k]).max() < 1e-3 * scale
    assert np.abs(im[k] + X.imag[k]).max() < 1e-3 * scale
    # bin 0 packs DC in re and Nyquist in im
    assert abs(re[0] - X.real[0]) < 1e-3 * scale and abs(im[0] - X.real[512]) < 1e-3 * scale
    # bin N/4 is never untangled: it is NOT the DFT bin
    assert x.sum() == 0
