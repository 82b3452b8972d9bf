"""The C ABI driven from a plain C99 host program (tests/host/abi_host_demo.c: no Python, no PyTorch in the process)
-- the stand-in for a cgo / JNI / FFI binding -- checked against the oracle on the signal the program wrote out."""
import os
import subprocess

import numpy as np
import pytest

from oracle import librosa_restated as lr
from oracle import pyin_restated as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_host_program(cuda, tmp_path):
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe, out = str(tmp_path / "abi_host_demo"), str(tmp_path / "out.bin")
    libdir = os.path.join(ROOT, "spev_tts_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(cuda_home, "include"), os.path.join(ROOT, "tests", "host", "abi_host_demo.c"),
                           "-o", exe, "-L" + libdir, "-lspev_b200", "-L" + os.path.join(cuda_home, "lib64"), "-lcudart", "-lm",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda_home, "lib64")])
    r = subprocess.run([exe, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, dtype=np.uint8)
    F = int(raw[:8].view(np.int64)[0])
    lens = [22050, 9000]
    assert F == sum(1 + n // 256 for n in lens)
    mel = raw[8: 8 + F * 80 * 4].view(np.float32).reshape(F, 80)
    states = raw[8 + F * 80 * 4: 8 + F * 80 * 4 + F * 4].view(np.int32)
    y = raw[8 + F * 80 * 4 + F * 4:].view(np.float32)
    starts, fo = [0, (lens[0] + 3) // 4 * 4], [0, 1 + lens[0] // 256, F]
    cfg = po.PyinConfig()
    for i, n in enumerate(lens):
        yi = y[starts[i]: starts[i] + n]
        assert np.abs(mel[fo[i]: fo[i + 1]] - lr.reference_logmel(yi)).max() <= 1e-4
        _, flag, _, st = po.pyin(yi, return_states=True)
        got = states[fo[i]: fo[i + 1]]
        assert np.mean((got < cfg.n_pitch_bins) == flag) >= 0.97
        both = (got < cfg.n_pitch_bins) & flag
        assert not both.any() or np.mean(np.abs(got[both] - st[both]) <= 1) >= 0.97
    v0 = states[: fo[1]] < cfg.n_pitch_bins
    assert v0[4:-4].all()                                    # the 150 Hz stack is voiced ...
    f0 = cfg.freqs[states[: fo[1]][4:-4]]
    assert np.abs(1200 * np.log2(f0 / 150.0)).max() <= 10.0  # ... at 150 Hz
    assert not (states[fo[1]:] < cfg.n_pitch_bins).any()     # and the noise item is unvoiced
