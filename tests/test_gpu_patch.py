"""GPU parity of the patch interpreter (K8, maximilian_b200/csrc/patch.cu) through the C ABI against the plain-C oracle and the
golden fixture made by the compiled reference.

Bars: patches whose stages are table look-ups and + - * / only are BIT-IDENTICAL; a patch with a stage that designs
coefficients or calls pow / atan / tan / cos on every sample (libdevice where the reference calls glibc) is asserted at
1e-9 relative + 1e-12 (north_star: 1e-5). Integer state (envelope flags, hold counts, delay ring indices, maxiEnvGen segment
and counters) is exact in both. The stereo bus sums voices in another (fixed) order: 1e-9 relative."""
import numpy as np
import pytest

import golden_checks as G
import patch_cases as PC
from maximilian_b200 import capi

pytestmark = pytest.mark.gpu

SLOTS = {1: 2, 2: 4, 3: 4, 4: 12, 5: 2, 6: 3, 7: 2, 8: 2, 10: 1, 11: 3}
INT_SLOTS = {2: (2, 3), 3: (2, 3), 4: (1, 2, 4), 10: (0,), 11: (0,)}          # integral members: exact whatever the patch


def _tables(port):
    g = G.load("tables")
    port.set_tables(g["sine"], g["transition"], float(g["sine_before"]), "port")
    capi.set_tables(g["sine"], g["transition"], float(g["sine_before"]))


def _cmp(got, ref, exact, what):
    if exact:
        assert np.array_equal(got, ref, equal_nan=True), f"{what}: not bit-identical, max abs err {np.nanmax(np.abs(got - ref))}"
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12, err_msg=what)


@pytest.mark.parametrize("V,B", [(333, 257), (4100, 64)])
@pytest.mark.parametrize("case", PC.cases(), ids=lambda c: c[0])
def test_patch_vs_oracle(port, case, V, B):
    name, d, params, inputs, exact, taps = case
    _tables(port)
    g = capi.Patch(d, V, max_frames=B, delay_taps=taps); o = port.Patch(d, V, delay_taps=taps, kind="port")
    for k, v in params(V, 21).items():
        g.set(k, v); o.set(k, v)
    for blk in range(3):
        ins = inputs(V, B, blk, 9)
        og, mg = g.process(B, ins, want_mix=True); oo, mo = o.process(B, ins, want_mix=True)
        _cmp(og, oo, exact, f"{name} blk{blk}")
        np.testing.assert_allclose(mg, mo, rtol=1e-9, atol=1e-10)
    for si, (op, kind, dst, src) in enumerate(d.stages):
        for sl in range(SLOTS.get(op, 0)):
            if op == 4 and sl == 3:
                continue
            _cmp(g.get_state(si, sl), o.get_state(si, sl), exact or sl in INT_SLOTS.get(op, ()), f"{name} stage {si} slot {sl}")
        if op in (10, 11):
            for v in range(0, V, 37):
                _cmp(g.ring(si, v, taps), o.ring(si, v, taps), exact, f"{name} ring {si}/{v}")


@pytest.mark.parametrize("case", PC.cases(), ids=lambda c: c[0])
def test_patch_golden(port, case):
    name, d, params, inputs, exact, taps = case
    _tables(port)
    gl = G.load("patches")
    V, B, NB = int(gl["V"]), int(gl["B"]), int(gl["NB"])
    p = capi.Patch(d, V, max_frames=B, delay_taps=taps)
    for k, v in params(V, 2468).items():
        p.set(k, v)
    for blk in range(NB):
        ins = {k: gl[f"{name}/in/{k}/{blk}"] for k in d.inputs}
        o, m = p.process(B, ins, want_mix=True)
        _cmp(o, gl[name + "/out"][blk], exact, f"{name} blk{blk}")
        np.testing.assert_allclose(m, gl[name + "/mix"][blk], rtol=1e-9, atol=1e-12)


def test_patch_rejects_bad_programs():
    from maximilian_b200.patchdef import PatchDef, R
    d = PatchDef(); d.stage("osc", d.K(100.0), kind="sinebuf", dst=R(0)); d.stage("out", R(0))
    ctx = capi.Context(0, 44100)                                    # a context that was never given the tables
    with pytest.raises(capi.MxbError):
        capi.Patch(d, 8, ctx=ctx, sample_rate=44100)
    d = PatchDef(); d.stage("delay", d.K(1.0), d.K(4.0), d.K(0.5), kind="dl", dst=R(0)); d.stage("out", R(0))
    with pytest.raises(capi.MxbError):
        capi.Patch(d, 8, delay_taps=0)
    d = PatchDef(); d.stages.append((99, 0, -1, [-1] * 8))
    with pytest.raises(capi.MxbError):
        capi.Patch(d, 8)
