"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
include/maxib200.h declares. No compute calls here (no GPU in this container); on a box without CUDA the
context constructor must fail loudly instead of falling back to anything."""
import os
import re

import pytest

from maximilian_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return build.build()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "maxib200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mxb_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol(built):
    assert os.path.exists(built)
    L = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/maxib200.h but not exported"
    assert set(declared) == set(capi.EXPORTS)
    assert L.mxb_version() == 100


def test_sass_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cuda_device_is_a_loud_error(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.MxbError):
        capi.Context(0, 48000)


def test_env_coeffs_host_helper_matches_oracle(built, port):
    import numpy as np
    ms = np.array([0.5, 1.0, 2.0, 17.3, 500.0])
    lib = port.load("port")
    for sr in (44100, 48000):
        a = capi.env_coeffs(0, ms, sr); am = capi.env_coeffs(1, ms, sr); d = capi.env_coeffs(2, ms, sr)
        for i, m in enumerate(ms):
            assert a[i] == lib.mxo_env_attack_coeff(m, sr)
            assert am[i] == lib.mxo_env_attack_ms_coeff(m, sr)
            assert d[i] == lib.mxo_env_decay_coeff(m, sr)
