"""``Vocoder`` drop-in: the Griffin-Lim fallback branch of the reference class
(``/root/reference/spev_real_metrics.py:709-736``, GL branch ``:725-733``).

``Vocoder(hifigan_dir).infer(mel)`` accepts what the reference accepts -- a torch tensor on
any device or a numpy array, ``[80,T]`` (``:785``) or ``[1,80,T]``
(``spev_embodied_core.py:250``, ``spev_temporal_policy.py:249``) log-mel in [-10, 2] -- and
returns a numpy float32 waveform of ``(T-1)*256`` samples per item, like the reference
(librosa broadcasts leading dims).  ``exp`` -> pinv/clip/sqrt -> Griffin-Lim all run on the GPU;
the only transfers are the mel in and the waveform out.

The neural (HiFi-GAN) branch of the reference class is out of scope; this class
always takes the Griffin-Lim path, i.e. it mirrors ``Vocoder`` with ``self.model is None``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import spectral

CONFIG = {"sr": 22050, "n_fft": 1024, "hop_length": 256, "n_mels": 80, "fmin": 0, "fmax": 8000}
# ^ spev_real_metrics.py:60-67


class Vocoder:
    def __init__(self, hifigan_dir: Optional[str] = None, *, n_iter: int = 32, device=None, nnls: str = "librosa"):
        # n_iter=32 is librosa's default, which is what the reference gets (it passes none);
        # BASELINE config 3 benchmarks n_iter=60.
        self.model = None
        self.n_iter = n_iter
        self.nnls = nnls      # "librosa": mel -> linear exactly as librosa solves it; "pinv": warm start only (no sync)
        self.device = torch.device(device) if device is not None else None

    def infer(self, mel, *, init_phase=None, random_state=None) -> np.ndarray:
        dev = self.device
        if not isinstance(mel, torch.Tensor):
            mel = torch.from_numpy(np.ascontiguousarray(mel, dtype=np.float32))
        if dev is None and mel.is_cuda:
            dev = mel.device
        y = spectral.mel_to_audio(mel, sr=CONFIG["sr"], n_fft=CONFIG["n_fft"],
                                  hop_length=CONFIG["hop_length"], fmin=CONFIG["fmin"],
                                  fmax=CONFIG["fmax"], n_iter=self.n_iter, is_log=True, nnls=self.nnls,
                                  init_phase=init_phase, random_state=random_state, device=dev)
        if not isinstance(y, torch.Tensor):
            return y
        # device -> pinned host at PCIe speed; the array keeps the pinned block alive (torch's host allocator
        # recycles it afterwards), so there is no extra host-side copy
        out = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
        out.copy_(y, non_blocking=True)
        torch.cuda.current_stream(y.device).synchronize()
        return out.numpy()
