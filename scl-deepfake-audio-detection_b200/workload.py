"""The benchmark's synthetic workload: seeded 64600-sample 16 kHz waveforms, the reference's default
RawBoost arguments (``/root/reference/main.py:258-298``) and parallel host-side plan drawing."""
from __future__ import annotations

import multiprocessing as mp
import os
from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np

from . import plans as _plans

UTT_LEN = 64600      # Dataset_for_eval.cut, asvspoof_2019_augall_3.py:152
SAMPLE_RATE = 16000

DEFAULT_ARGS = dict(
    algo=5, nBands=5, minF=20, maxF=8000, minBW=100, maxBW=1000, minCoeff=10, maxCoeff=100, minG=0, maxG=0,
    minBiasLinNonLin=5, maxBiasLinNonLin=20, N_f=5, P=10, g_sd=2, SNRmin=10, SNRmax=40,
)


def default_args(**overrides) -> SimpleNamespace:
    kw = dict(DEFAULT_ARGS)
    kw.update(overrides)
    return SimpleNamespace(**kw)


def synth_utterance(u: int, length: int = UTT_LEN, loud: bool = False) -> np.ndarray:
    """Utterance ``u``: speech-level gaussian clipped to [-1, 1] (normWav mostly idle) or the loud uniform variant."""
    rs = np.random.RandomState(20240000 + u)
    if loud:
        return (0.9 * rs.uniform(-1, 1, length)).astype(np.float32)
    return np.clip(0.1 * rs.standard_normal(length), -1, 1).astype(np.float32)


def seed_for(u: int) -> int:
    """Seed of the global numpy stream set immediately before utterance ``u`` (1234 = main.py:239)."""
    return (1234 + u) % 2 ** 32


def synth_batch(first: int, count: int, length: int = UTT_LEN, ld: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
    """[count, ld] float32 batch of utterances first..first+count-1 (odd ones loud)."""
    ld = ld or _plans.padded_ld(length)
    if out is None:
        out = np.zeros((count, ld), dtype=np.float32)
    for i in range(count):
        out[i, :length] = synth_utterance(first + i, length, bool((first + i) % 2))
    return out


# ---- parallel plan drawing: utterances are seeded independently, so workers can draw any subset -------------
def _draw_chunk(job):
    lengths, sr, args_kw, algo, seeds = job
    args = SimpleNamespace(**args_kw)
    out = []
    for n, s in zip(lengths, seeds):
        np.random.seed(int(s))
        out.append(_plans.draw_for_algo(int(n), sr, args, algo))
    return out


class PlanPool:
    """A pool of host processes that draw per-utterance plans with the reference's numpy calls.

    Start it BEFORE CUDA is initialised in the parent (the workers are forked and never touch CUDA)."""

    def __init__(self, workers: Optional[int] = None):
        self.workers = max(1, workers or (os.cpu_count() or 1))
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None

    def draw_batch(self, lengths: Sequence[int], sr, args, algo: int, seeds: Sequence[int], ld: Optional[int] = None):
        n = len(lengths)
        if self.pool is None or n < 2 * self.workers:
            plans = _draw_chunk((list(lengths), sr, vars(args), algo, list(seeds)))
        else:
            step = max(1, (n + 4 * self.workers - 1) // (4 * self.workers))
            jobs = [(list(lengths[i:i + step]), sr, vars(args), algo, list(seeds[i:i + step])) for i in range(0, n, step)]
            plans = [p for chunk in self.pool.map(_draw_chunk, jobs) for p in chunk]
        return _plans.pack(plans, ld=ld)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
