"""Reverb view on the device: the arithmetic of ``ReverbAugmentor.transform``
(``/root/reference/datautils/audio_augmentor/reverb.py:33-44``) -- a full convolution of the waveform with a room impulse
response followed by peak normalisation -- on the FIR-bank kernel (long filters run as 512-tap segments) and the one-kernel
``normWav``.

    reverberate = np.convolve(data, rir_data)            # length len(data) + len(rir) - 1
    reverberate /= np.max(np.abs(reverberate))

What stays with the caller, exactly as in the reference: choosing and loading the RIR file (``random.choice`` over the
corpus, ``librosa.load``) and the 16-bit pydub round trip of the augmentor framework. There is no CPU fallback.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .engine import Engine, default_engine
from .plans import padded_ld


def reverb_batch(eng: Engine, waves: Sequence[np.ndarray], rirs: Sequence[np.ndarray]):
    """``normWav(np.convolve(wave_u, rir_u), 1)`` for every utterance; returns ([B, ld] float32 device tensor, int32 lengths).

    filterFIR computes y[n] = sum_k b[k] x[n + (K+1)//2 - k] (RawBoost.py:51-56); feeding it the waveform delayed by
    (K+1)//2 samples and zero-extended to len + K - 1 turns that into the full convolution sum_k b[k] x[n - k]."""
    if len(waves) != len(rirs):
        raise ValueError("one impulse response per waveform")
    B = len(waves)
    rirs = [np.asarray(r, dtype=np.float32).reshape(-1) for r in rirs]
    out_len = np.array([w.shape[0] + r.shape[0] - 1 if w.shape[0] and r.shape[0] else 0 for w, r in zip(waves, rirs)], dtype=np.int32)
    shift = [(r.shape[0] + 1) // 2 for r in rirs]
    in_len = np.array([n + s if n else 0 for n, s in zip(out_len, shift)], dtype=np.int32)
    ld = padded_ld(int(in_len.max()) if B else 0)
    host = np.zeros((B, ld), dtype=np.float32)
    for u, (w, s) in enumerate(zip(waves, shift)):
        host[u, s:s + w.shape[0]] = np.asarray(w, dtype=np.float32)
    x = torch.from_numpy(host).to(eng.device)
    taps = torch.from_numpy(np.concatenate(rirs) if B else np.zeros(0, np.float32)).to(eng.device)
    off = torch.from_numpy(np.concatenate([[0], np.cumsum([r.shape[0] for r in rirs])]).astype(np.int32)).to(eng.device)
    # the kernel takes one length per row for both input and output: run over the delayed input's length; the outputs beyond
    # len + K - 1 are exact zeros (no tap reaches a sample there), so they do not disturb the peak
    ln_in = torch.from_numpy(in_len).to(eng.device)
    y = eng.filter_fir(x, ln_in, taps, off)
    return eng.normwav(y, ln_in, True), torch.from_numpy(out_len).to(eng.device)


def reverb_convolve(data, rir_data) -> np.ndarray:
    """One utterance: float32 array of length ``len(data) + len(rir_data) - 1``, peak-normalised like the reference."""
    data = np.asarray(data).reshape(-1)
    rir_data = np.asarray(rir_data).reshape(-1)
    if data.shape[0] == 0 or rir_data.shape[0] == 0:
        raise ValueError("reverb_convolve needs a non-empty waveform and impulse response (np.convolve raises too)")
    y, ln = reverb_batch(default_engine(), [data], [rir_data])
    return y[0, :int(ln[0])].cpu().numpy()
