"""The C++ host layer (include/maximilian_b200.hpp): a reference-style patch compiles against it (CPU check) and,
on the GPU, its audio-callback output equals the oracle running the same chain per sample."""
import os
import subprocess

import numpy as np
import pytest

from maximilian_b200 import build
from maximilian_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "patch_poly.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "patch_poly")


def compile_patch():
    lib = build.build()
    libdir = os.path.dirname(lib)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", libdir, "-lmaxib200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return EXE


def test_patch_compiles_against_the_dropin_header():
    exe = compile_patch()
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_patch_through_routing_matches_oracle(port, tmp_path):
    exe = compile_patch()
    V, B, NB, cap = 300, 256, 3, 512
    p = W.voice_params(V, seed=31, delay_size=cap, ragged_delay=True)
    gates = [W.gate(V, B, 4 * k) for k in range(NB + 1)]
    with open(tmp_path / "params.bin", "wb") as f:
        for k in ("freq", "phase", "cutoff", "res_svf", "attack_ms", "decay_ms", "env_sustain", "release_ms",
                  "delay_size", "delay_feedback", "pan", "gain"):
            f.write(np.ascontiguousarray(p[k], dtype=np.float64).tobytes())
        for on, off in gates:
            f.write(on.astype(np.int32).tobytes()); f.write(off.astype(np.int32).tobytes())
    r = subprocess.run([exe, str(tmp_path / "params.bin"), str(tmp_path / "out.bin"), str(V), str(B), str(NB)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(tmp_path / "out.bin", dtype=np.float64)
    mix = raw[:NB * B * 2].reshape(NB, B, 2)
    voices = raw[NB * B * 2:].reshape(B, V)

    o = port.Bank(V, osc="saw", filt="svf", env=True, delay=True, delay_capacity=cap)
    W.configure_bank(o, "svf", p, env=True, delay=True)
    for k in range(NB):
        _, mo = o.process(B, gates[k][0], gates[k][1], want_out=False, want_mix=True)
        np.testing.assert_allclose(mix[k], mo, rtol=1e-9, atol=1e-11)
    oo, _ = o.process(B, gates[NB][0], gates[NB][1])
    assert np.array_equal(voices, oo)
