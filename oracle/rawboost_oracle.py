"""CPU oracle for the RawBoost hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or as the timed CPU
baseline. The product path (``scl-deepfake-audio-detection_b200/``) never imports it and has no
CPU fallback.

It is a numpy/scipy restatement (float64, same third-party primitives and the same order of
draws on the process-global legacy ``np.random`` stream) of the reference algorithm in
``/root/reference/datautils/RawBoost.py:14-97`` and of the 9-way dispatcher
``/root/reference/datautils/asvspoof_2019_augall_3.py:377-439``. The restatement is organised as
*draw a plan* / *apply a plan* pairs, because the split is what the CUDA side consumes: every
random draw of an operator happens before any of its arithmetic, so draw-then-apply leaves the
global RNG stream bit-identical to the reference's interleaved order.

Parity pin: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, imported unmodified in the build
container by ``oracle/make_golden.py``; the resulting fixtures live in ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks this file against them (bit-exact for plans / indices,
<=1e-12 for float64 waveforms).

Third-party arithmetic the reference leans on (not vendored by it): numpy (unpinned) for
``random.uniform/permutation/rand/normal``, ``convolve``, ``power``; scipy.signal (pinned 1.7.3 in
``00_envsetup.sh:32``; 1.18.1 here) for ``firwin``, ``freqz``, ``lfilter``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
from scipy import signal

__all__ = [
    "rand_range", "norm_wav", "draw_notch_taps", "filter_fir", "filter_fir_closed_form",
    "LnLPlan", "ISDPlan", "SSIPlan", "draw_lnl_plan", "apply_lnl", "draw_isd_plan", "apply_isd",
    "draw_ssi_plan", "apply_ssi", "lnl", "isd", "ssi", "process", "DEFAULT_ARGS", "make_args",
    "synth_utterance", "seed_for", "dataset_item", "overscale_utterance", "corpus_wave", "CORPUS_IDS", "CORPUS_VOCODERS",
]


# --------------------------------------------------------------------------------------------
# scalars and the two pointwise helpers
# --------------------------------------------------------------------------------------------
def rand_range(lo, hi, integer):
    """One size-(1,) uniform on the global legacy stream (RawBoost.py:14-18).

    Non-integer results stay shape-(1,) float64 arrays; integer results are truncated toward
    zero. ``lo > hi`` is legal (numpy computes ``lo + (hi-lo)*u``)."""
    draw = np.random.uniform(low=lo, high=hi, size=(1,))
    return int(draw[0]) if integer else draw


def norm_wav(x, always):
    """Peak normalisation (RawBoost.py:20-25): divide by max|x| if ``always`` or the peak is > 1."""
    peak = np.amax(np.abs(x))
    if always or peak > 1:
        return x / peak
    return x


# --------------------------------------------------------------------------------------------
# notch-filter cascade (genNotchCoeffs) and the delay-compensated FIR (filterFIR)
# --------------------------------------------------------------------------------------------
def draw_notch_taps(nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff, minG, maxG, fs):
    """Cascade of ``nBands`` Hamming band-stop FIRs with a random gain (RawBoost.py:28-48).

    Stream order: (fc, bw, c) per band, then G -> 3*nBands+1 uniforms. ``c`` is forced odd.
    The cascade is peak-normalised on scipy's default 512-point ``freqz`` grid.
    Returns float64 taps of length sum(c) - (nBands-1)."""
    taps = np.ones(1)
    nyq = fs / 2
    for _ in range(nBands):
        centre = float(rand_range(minF, maxF, 0)[0])
        width = float(rand_range(minBW, maxBW, 0)[0])
        ntaps = rand_range(minCoeff, maxCoeff, 1)
        if ntaps % 2 == 0:
            ntaps += 1
        lo = centre - width / 2
        hi = centre + width / 2
        if lo <= 0:
            lo = 1 / 1000
        if hi >= nyq:
            hi = nyq - 1 / 1000
        stage = signal.firwin(ntaps, [lo, hi], window="hamming", fs=fs)
        taps = np.convolve(stage, taps)
    gain_db = float(rand_range(minG, maxG, 0)[0])
    _, resp = signal.freqz(taps, 1, fs=fs)
    return (10.0 ** (gain_db / 20)) * taps / np.amax(np.abs(resp))


def filter_fir(x, taps):
    """'Same'-length FIR with the reference's delay compensation (RawBoost.py:51-56).

    Pads K+1 zeros at the end, runs the causal filter and drops (K+1)/2 samples at each end."""
    n_ext = taps.shape[0] + 1
    padded = np.concatenate([np.asarray(x), np.zeros(n_ext, dtype=np.asarray(x).dtype)])
    causal = signal.lfilter(taps, 1, padded)
    return causal[int(n_ext / 2): int(causal.shape[0] - n_ext / 2)]


def filter_fir_closed_form(x, taps):
    """y[n] = sum_k taps[k] * x[n + (K+1)//2 - k], x == 0 outside [0, L)  (SURVEY.md 8a-4).

    The formula the CUDA kernel implements; ``tests`` check it equals :func:`filter_fir`."""
    x = np.asarray(x, dtype=np.float64)
    taps = np.asarray(taps, dtype=np.float64)
    L, K = x.shape[0], taps.shape[0]
    shift = (K + 1) // 2
    full = np.convolve(x, taps)  # full[m] = sum_k taps[k] x[m-k], length L+K-1
    out = np.zeros(L)
    hi = min(L, L + K - 1 - shift)
    out[:hi] = full[shift: shift + hi]
    return out


# --------------------------------------------------------------------------------------------
# plans
# --------------------------------------------------------------------------------------------
@dataclass
class LnLPlan:
    """Taps of the N_f notch cascades; filter i is applied to x**(i+1)."""
    taps: List[np.ndarray] = field(default_factory=list)


@dataclass
class ISDPlan:
    """Impulse positions (int64, unique) and their signed gains f_r in (-1, 1)."""
    beta: float = 0.0
    idx: np.ndarray = None
    f_r: np.ndarray = None


@dataclass
class SSIPlan:
    """White noise (float64, length L), one notch cascade and the SNR in dB."""
    noise: np.ndarray = None
    taps: np.ndarray = None
    snr_db: float = 0.0


def draw_lnl_plan(N_f, nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff, minG, maxG,
                  minBiasLinNonLin, maxBiasLinNonLin, fs) -> LnLPlan:
    """RNG part of LnL (RawBoost.py:61-65): the gain range drops once, at order 2, and stays."""
    plan = LnLPlan()
    g_lo, g_hi = minG, maxG
    for order in range(N_f):
        if order == 1:
            g_lo = g_lo - minBiasLinNonLin
            g_hi = g_hi - maxBiasLinNonLin
        plan.taps.append(draw_notch_taps(nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff,
                                         g_lo, g_hi, fs))
    return plan


def apply_lnl(x, plan: LnLPlan):
    """Arithmetic part of LnL (RawBoost.py:60,66-69): powers in x's dtype, filters and the sum in
    float64, mean removal, conditional peak normalisation."""
    acc = np.zeros(x.shape[0])
    for order, taps in enumerate(plan.taps):
        acc = acc + filter_fir(np.power(x, order + 1), taps)
    acc = acc - np.mean(acc)
    return norm_wav(acc, 0)


def draw_isd_plan(length, P) -> ISDPlan:
    """RNG part of ISD (RawBoost.py:74,78-80): beta, a full permutation of [0,L), two rand(n)."""
    beta = float(rand_range(0, P, 0)[0])
    count = int(length * (beta / 100))
    idx = np.random.permutation(length)[:count]
    f_r = np.multiply((2 * np.random.rand(idx.shape[0])) - 1, (2 * np.random.rand(idx.shape[0])) - 1)
    return ISDPlan(beta=beta, idx=idx, f_r=f_r)


def apply_isd(x, plan: ISDPlan, g_sd):
    """Arithmetic part of ISD (RawBoost.py:76,81-84). Output keeps x's dtype; x is not mutated."""
    out = np.array(x, copy=True)
    hit = x[plan.idx]
    out[plan.idx] = hit + g_sd * hit * plan.f_r
    return norm_wav(out, 0)


def draw_ssi_plan(length, SNRmin, SNRmax, nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff,
                  minG, maxG, fs) -> SSIPlan:
    """RNG part of SSI (RawBoost.py:90-91,94): L normals, one notch cascade, the SNR."""
    noise = np.random.normal(0, 1, length)
    taps = draw_notch_taps(nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff, minG, maxG, fs)
    snr = float(rand_range(SNRmin, SNRmax, 0)[0])
    return SSIPlan(noise=noise, taps=taps, snr_db=snr)


def apply_ssi(x, plan: SSIPlan):
    """Arithmetic part of SSI (RawBoost.py:92-97): colour, peak-normalise, scale to the SNR, add.
    No final normalisation -- the result may exceed 1."""
    coloured = norm_wav(filter_fir(plan.noise, plan.taps), 1)
    coloured = coloured / np.linalg.norm(coloured, 2) * np.linalg.norm(x, 2) / 10.0 ** (0.05 * plan.snr_db)
    return x + coloured


# --------------------------------------------------------------------------------------------
# the three operators and the dispatcher, reference argument order
# --------------------------------------------------------------------------------------------
def lnl(x, N_f, nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff, minG, maxG,
        minBiasLinNonLin, maxBiasLinNonLin, fs):
    return apply_lnl(x, draw_lnl_plan(N_f, nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff,
                                      minG, maxG, minBiasLinNonLin, maxBiasLinNonLin, fs))


def isd(x, P, g_sd):
    return apply_isd(x, draw_isd_plan(x.shape[0], P), g_sd)


def ssi(x, SNRmin, SNRmax, nBands, minF, maxF, minBW, maxBW, minCoeff, maxCoeff, minG, maxG, fs):
    return apply_ssi(x, draw_ssi_plan(x.shape[0], SNRmin, SNRmax, nBands, minF, maxF, minBW, maxBW,
                                      minCoeff, maxCoeff, minG, maxG, fs))


def process(feature, sr, args, algo):
    """The 9-way switch (asvspoof_2019_augall_3.py:377-439): 1 LnL, 2 ISD, 3 SSI, 4 LnL>ISD>SSI,
    5 LnL>ISD, 6 LnL>SSI, 7 ISD>SSI, 8 normWav(LnL+ISD), anything else identity (same object)."""
    a = args

    def _lnl(v):
        return lnl(v, a.N_f, a.nBands, a.minF, a.maxF, a.minBW, a.maxBW, a.minCoeff, a.maxCoeff,
                   a.minG, a.maxG, a.minBiasLinNonLin, a.maxBiasLinNonLin, sr)

    def _isd(v):
        return isd(v, a.P, a.g_sd)

    def _ssi(v):
        return ssi(v, a.SNRmin, a.SNRmax, a.nBands, a.minF, a.maxF, a.minBW, a.maxBW, a.minCoeff,
                   a.maxCoeff, a.minG, a.maxG, sr)

    if algo == 1:
        return _lnl(feature)
    if algo == 2:
        return _isd(feature)
    if algo == 3:
        return _ssi(feature)
    if algo == 4:
        return _ssi(_isd(_lnl(feature)))
    if algo == 5:
        return _isd(_lnl(feature))
    if algo == 6:
        return _ssi(_lnl(feature))
    if algo == 7:
        return _ssi(_isd(feature))
    if algo == 8:
        first = _lnl(feature)
        second = _isd(feature)
        return norm_wav(first + second, 0)
    return feature


# --------------------------------------------------------------------------------------------
# the benchmark's synthetic workload -- shared by tests and bench.py
# --------------------------------------------------------------------------------------------
DEFAULT_ARGS = dict(  # /root/reference/main.py:258-298
    algo=5, nBands=5, minF=20, maxF=8000, minBW=100, maxBW=1000, minCoeff=10, maxCoeff=100,
    minG=0, maxG=0, minBiasLinNonLin=5, maxBiasLinNonLin=20, N_f=5, P=10, g_sd=2,
    SNRmin=10, SNRmax=40,
)


# --------------------------------------------------------------------------------------------------------
# the step right after the path: length unification + one shared crop over the views of an item
# (/root/reference/core_scripts/data_io/wav_augmentation.py:209-282), restated as an index map
# --------------------------------------------------------------------------------------------------------
def multiview_crop_plan(first_len, length, random_trim_nosil, repeat_pad):
    """(start, out_len, wrap) of ``batch_pad_for_multiview``. Draws ``np.random.rand()`` exactly when the reference does
    (lines 256 / 273: only when the first view is at least ``length`` long and ``random_trim_nosil`` is set)."""
    new_len = int(first_len)
    if new_len < length:
        return (0, length, True) if repeat_pad else (0, new_len, False)
    start = int(np.random.rand() * (new_len - length)) if random_trim_nosil else 0
    return start, length, False


def multiview_gather(views, first_len, start, out_len, wrap, repeat_pad):
    """out[k, v] = adjusted_v[(start + k) mod first_len if wrap else start + k], where adjusted_v is view v cut to
    ``first_len`` samples or extended to it by zeros / by repetition (``_ad_length``, lines 229-241)."""
    cols = []
    for v in views:
        v = np.asarray(v).reshape(-1)
        m = start + np.arange(out_len)
        if wrap:
            m = m % first_len
        if repeat_pad:
            col = v[m % v.shape[0]]
        else:
            col = np.where(m < v.shape[0], v[np.minimum(m, v.shape[0] - 1)], 0)
        cols.append(col)
    return np.stack(cols, axis=1)


def batch_pad_for_multiview(input_data_batch_, wav_samp_rate, length, random_trim_nosil=False, repeat_pad=False):
    """Same signature and result as the reference function: list of (out_len, 1) arrays."""
    first_len = input_data_batch_[0].shape[0]
    start, out_len, wrap = multiview_crop_plan(first_len, length, random_trim_nosil, repeat_pad)
    out = multiview_gather([x[:, 0] for x in input_data_batch_], first_len, start, out_len, wrap, repeat_pad)
    return [out[:, v:v + 1] for v in range(out.shape[1])]


def reverb_convolve(data, rir_data):
    """Arithmetic of ``ReverbAugmentor.transform`` (/root/reference/datautils/audio_augmentor/reverb.py:39-42)."""
    reverberate = np.convolve(np.asarray(data, dtype=np.float64), np.asarray(rir_data, dtype=np.float64))
    reverberate /= np.max(np.abs(reverberate))
    return reverberate


def dataset_item(idx, list_ids, load_audio, args, vocoders, num_additional_real, trim_length, sr=16000, repeat_pad=True,
                 bonafide_dir="/data/bonafide", vocoded_dir="/data/vocoded"):
    """``Dataset_for.__getitem__`` with ``augmentation_methods == ['RawBoost12']``, ``online_aug`` on
    (/root/reference/datautils/asvspoof_2019_augall_3.py:103-146). Draw order on the global stream: RawBoost (algo 5) on
    every vocoded copy, RawBoost on the anchor, ``np.random.choice`` of the additional bona fide utterances, the shared crop.
    View order: anchor, augmented anchor, additional, vocoded, augmented vocoded. Returns (utt id, [length, V] float32,
    labels float32): 1 for the anchor / positives, 0 for everything vocoded."""
    import os
    anchor_path = os.path.join(bonafide_dir, list_ids[idx])
    anchor = load_audio(anchor_path)
    vocoded, aug_vocoded = [], []
    for v in vocoders:
        w = load_audio(os.path.join(vocoded_dir, v + "_" + list_ids[idx]))
        vocoded.append(np.expand_dims(w, 1))
        aug_vocoded.append(np.expand_dims(process(w, sr, args, 5), 1))
    augmented = [np.expand_dims(process(anchor, sr, args, 5), 1)]
    others = list(range(len(list_ids)))
    others.remove(idx)
    extra_idx = np.random.choice(others, num_additional_real, replace=False)
    extra = [np.expand_dims(load_audio(os.path.join(bonafide_dir, list_ids[i])), 1) for i in extra_idx]
    views = [np.expand_dims(anchor, 1)] + augmented + extra + vocoded + aug_vocoded
    out = batch_pad_for_multiview(views, sr, trim_length, random_trim_nosil=True, repeat_pad=repeat_pad)
    data = np.concatenate(out, axis=1).astype(np.float32)
    label = np.array([1] * (len(augmented) + len(extra) + 1) + [0] * (2 * len(vocoders)), dtype=np.float32)
    return list_ids[idx], data, label


def overscale_utterance(u: int, length: int) -> np.ndarray:
    """Input above full scale (peak ~2.5): the case in which it matters that ISD sees the raw x before normWav."""
    return (2.5 * np.random.RandomState(777 + u).uniform(-1, 1, length)).astype(np.float32)


CORPUS_IDS = [f"utt{k}.wav" for k in range(6)]
CORPUS_VOCODERS = ["hifigan", "hn-sinc-nsf-hifi", "waveglow"]


def corpus_wave(path: str) -> np.ndarray:
    """The synthetic stand-in for ``librosa.load`` used by the ``__getitem__`` fixtures: ``<dir>/utt<k>.wav`` and
    ``<dir>/<vocoder>_utt<k>.wav`` map to seeded waveforms of slightly different lengths."""
    import os
    name = os.path.basename(path)
    for vi, v in enumerate(CORPUS_VOCODERS):
        if name.startswith(v + "_"):
            k = CORPUS_IDS.index(name[len(v) + 1:])
            return synth_utterance(900 + 10 * k + vi + 1, 15000 + 111 * k + 13 * vi, bool(vi % 2))
    k = CORPUS_IDS.index(name)
    return synth_utterance(900 + 10 * k, 15000 + 111 * k, False)


class _Args:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_args(**overrides):
    kw = dict(DEFAULT_ARGS)
    kw.update(overrides)
    return _Args(**kw)


def synth_utterance(u: int, length: int = 64600, loud: bool = False) -> np.ndarray:
    """Utterance ``u`` of the synthetic workload: speech-level gaussian (normWav mostly idle) or
    the loud uniform variant (normWav always fires). float32, as ``librosa.load`` would give."""
    rs = np.random.RandomState(20240000 + u)
    if loud:
        return (0.9 * rs.uniform(-1, 1, length)).astype(np.float32)
    return np.clip(0.1 * rs.standard_normal(length), -1, 1).astype(np.float32)


def seed_for(u: int) -> int:
    """Global-stream seed set immediately before utterance ``u`` on both sides (1234 = main.py:239)."""
    return (1234 + u) % 2 ** 32
