"""The step right after RawBoost for the views of a training item, kept on the device.

* :func:`batch_pad_for_multiview` -- drop-in for ``core_scripts/data_io/wav_augmentation.py:209-282`` (same signature,
  same ``np.random.rand()`` draw in the same place, same result): every view is cut / zero-extended / tiled to the length
  of view 0 and one shared crop is taken.
* :func:`crop_plan` -- the host part alone: ``(start, out_len)`` of that crop; the global numpy stream is left exactly
  where the reference leaves it.
* :func:`assemble` -- the device part alone, batched over items: views already on the device (e.g. straight out of
  :meth:`engine.Engine.process`) go to ``[G, length, V]`` (the Dataset's ``batch_data``, asvspoof_2019_augall_3.py:138-142)
  or ``[G, V, length]`` (what the model consumes after main.py:57-60) without visiting the host.
* :func:`item_views` -- one call for whole items in the loader's order (RawBoost12 on the vocoded copies, then on the
  anchor, then the crop: asvspoof_2019_augall_3.py:109-138).

Integer work (crop start, index map) is bit-exact; the samples are copies of the views' float32 values.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, plans as _plans
from .engine import Engine, default_engine

LAYOUT_ITEM = 0    # [G, length, V]  (np.concatenate(views, axis=1) of the reference)
LAYOUT_MODEL = 1   # [G, V, length]  (batch_x after the transpose in main.py:57-60)


def crop_plan(first_len: int, length: int, random_trim_nosil: bool = False, repeat_pad: bool = False) -> Tuple[int, int]:
    """``(start, out_len)`` of the shared crop. ``np.random.rand()`` is consumed iff the reference consumes it: only when
    view 0 has at least ``length`` samples and ``random_trim_nosil`` is set (wav_augmentation.py:256, 273)."""
    first_len, length = int(first_len), int(length)
    if first_len < length:
        return 0, (length if repeat_pad else first_len)
    start = int(np.random.rand() * (first_len - length)) if random_trim_nosil else 0
    return start, length


def assemble(eng: Engine, views: torch.Tensor, lengths: torch.Tensor, V: int, starts, length: int, repeat_pad: bool,
             layout: int = LAYOUT_MODEL, out: Optional[torch.Tensor] = None):
    """views: [G*V, ld] float32 on the device (view v of item g in row g*V+v), lengths: int32 [G*V], starts: G crop starts.
    Returns ``(out, out_len)``: out is [G, length, V] or [G, V, length]; out_len int32 [G] = samples written per view.
    Asynchronous on torch's current stream."""
    if views.dtype != torch.float32 or views.dim() != 2 or not views.is_contiguous() or views.device != eng.device:
        raise ValueError("views must be a contiguous [G*V, ld] float32 tensor on the engine's device")
    rows, ld = views.shape
    if V <= 0 or rows % V:
        raise ValueError("rows of `views` must be a multiple of V")
    G = rows // V
    if lengths.dtype != torch.int32 or lengths.numel() != rows or lengths.device != eng.device:
        raise ValueError("lengths must be int32 [G*V] on the engine's device")
    if not torch.is_tensor(starts):
        starts = torch.tensor(np.asarray(starts, dtype=np.int32), device=eng.device)
    if starts.dtype != torch.int32 or starts.numel() != G:
        raise ValueError("one int32 crop start per item")
    shape = (G, length, V) if layout == LAYOUT_ITEM else (G, V, length)
    if out is None:
        out = torch.zeros(shape, dtype=torch.float32, device=eng.device)
    out_len = torch.empty(G, dtype=torch.int32, device=eng.device)
    stream = torch.cuda.current_stream(eng.device).cuda_stream
    with torch.cuda.device(eng.device):
        rc = eng.lib.rb_multiview_assemble(C.c_void_p(views.data_ptr()), C.c_void_p(lengths.data_ptr()), G, V, ld,
                                           C.c_void_p(starts.data_ptr()), int(length), int(bool(repeat_pad)), int(layout),
                                           C.c_void_p(out.data_ptr()), C.c_void_p(out_len.data_ptr()), C.c_void_p(stream))
    _lib.check(rc, "rb_multiview_assemble")
    return out, out_len


def assemble_ex(eng: Engine, a: torch.Tensor, b: Optional[torch.Tensor], view_row: torch.Tensor, lengths: torch.Tensor, V: int,
                starts, length: int, repeat_pad: bool, layout: int = LAYOUT_MODEL, out: Optional[torch.Tensor] = None,
                view_label: Optional[torch.Tensor] = None):
    """:func:`assemble` with the views read where they already are (``rb_multiview_assemble_ex``): view v of item g is row
    ``r = view_row[g*V+v]`` of ``a`` (r >= 0) or row ``-1-r`` of ``b`` (r < 0) -- typically the original waveforms and their
    RawBoost results, two [R, ld] tensors sharing ``lengths`` [R]. No regrouping copy. With ``view_label`` (float32 [V]) the
    item's label vector (asvspoof_2019_augall_3.py:143-146) is produced too. Returns ``(out, out_len, labels or None)``."""
    for t in (a, b):
        if t is not None and (t.dtype != torch.float32 or t.dim() != 2 or not t.is_contiguous() or t.device != eng.device):
            raise ValueError("sources must be contiguous [R, ld] float32 tensors on the engine's device")
    if b is not None and b.shape != a.shape:
        raise ValueError("both sources must have the same shape")
    R, ld = a.shape
    if view_row.dtype != torch.int32 or view_row.device != eng.device or view_row.numel() % V:
        raise ValueError("view_row must be an int32 device tensor with G*V entries")
    G = view_row.numel() // V
    if lengths.dtype != torch.int32 or lengths.numel() != R or lengths.device != eng.device:
        raise ValueError("lengths must be int32 [R] on the engine's device")
    if not torch.is_tensor(starts):
        starts = torch.tensor(np.asarray(starts, dtype=np.int32), device=eng.device)
    if starts.dtype != torch.int32 or starts.numel() != G:
        raise ValueError("one int32 crop start per item")
    shape = (G, length, V) if layout == LAYOUT_ITEM else (G, V, length)
    if out is None:
        out = torch.zeros(shape, dtype=torch.float32, device=eng.device)
    out_len = torch.empty(G, dtype=torch.int32, device=eng.device)
    labels = None
    if view_label is not None:
        view_label = view_label.to(device=eng.device, dtype=torch.float32).contiguous()
        labels = torch.empty((G, V), dtype=torch.float32, device=eng.device)
    stream = torch.cuda.current_stream(eng.device).cuda_stream
    with torch.cuda.device(eng.device):
        rc = eng.lib.rb_multiview_assemble_ex(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()) if b is not None else None,
                                              C.c_void_p(view_row.data_ptr()), C.c_void_p(lengths.data_ptr()), G, V, ld,
                                              C.c_void_p(starts.data_ptr()), int(length), int(bool(repeat_pad)), int(layout),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(out_len.data_ptr()),
                                              C.c_void_p(view_label.data_ptr()) if view_label is not None else None,
                                              C.c_void_p(labels.data_ptr()) if labels is not None else None, C.c_void_p(stream))
    _lib.check(rc, "rb_multiview_assemble_ex")
    return out, out_len, labels


def batch_pad_for_multiview(input_data_batch_, wav_samp_rate, length, random_trim_nosil=False, repeat_pad=False):
    """Drop-in for the reference function: list of (len_v, 1) arrays in, list of (out_len, 1) float32 arrays out."""
    eng = default_engine()
    waves = [np.asarray(x, dtype=np.float32).reshape(-1) for x in input_data_batch_]
    start, out_len = crop_plan(waves[0].shape[0], length, random_trim_nosil, repeat_pad)
    if out_len == 0:
        return [np.zeros((0, 1), dtype=np.float32) for _ in waves]
    xd, ln = eng.pack_waveforms(waves)
    out, _ = assemble(eng, xd, ln, len(waves), [start], int(length), bool(repeat_pad), LAYOUT_MODEL)
    host = out[0].cpu().numpy()
    return [host[v, :out_len].reshape(out_len, 1) for v in range(len(waves))]


def item_view_rows(G: int, nvoc: int, extra_rows: Optional[np.ndarray] = None) -> np.ndarray:
    """Row table of :func:`assemble_ex` for G items whose waveforms sit in the loader's draw order -- per item ``nvoc`` vocoded
    copies, then the anchor (asvspoof_2019_augall_3.py:109-124) -- in one [G*(nvoc+1), ld] input tensor ``x`` with the RawBoost
    results in ``y``. View order of the Dataset (line 133): anchor, augmented anchor, [additional bona fide rows], vocoded,
    augmented vocoded. ``extra_rows``: int [G, n_add] rows of ``x`` holding each item's additional bona fide utterances."""
    per = nvoc + 1
    base = np.arange(G, dtype=np.int64)[:, None] * per
    voc = base + np.arange(nvoc)[None, :]
    anchor = base + nvoc
    cols = [anchor, -1 - anchor]
    if extra_rows is not None and np.size(extra_rows):
        cols.append(np.asarray(extra_rows, dtype=np.int64).reshape(G, -1))
    cols += [voc, -1 - voc]
    return np.concatenate(cols, axis=1).astype(np.int32)


def item_labels(nvoc: int, n_add: int = 0, n_aug: int = 1) -> np.ndarray:
    """Label vector of one item (asvspoof_2019_augall_3.py:143-146): 1 for the anchor and its positives, 0 for everything vocoded."""
    return np.array([1.0] * (n_aug + n_add + 1) + [0.0] * (2 * nvoc), dtype=np.float32)


def item_views(eng: Engine, items: Sequence[Tuple[np.ndarray, Sequence[np.ndarray]]], args, sr: int, trim_length: int,
               repeat_pad: bool = True, random_trim_nosil: bool = True, layout: int = LAYOUT_MODEL, with_labels: bool = False):
    """Views of whole items, RawBoost and assembly on the device, random draws in the loader's order.

    ``items``: ``(anchor, [vocoded copies])`` float32 waveforms. Per item, as Dataset_for.__getitem__ with
    ``augmentation_methods[0] == 'RawBoost12'`` does (asvspoof_2019_augall_3.py:109-138): RawBoost (algo 5) is drawn for
    each vocoded copy, then for the anchor, then the shared crop. View order: anchor, augmented anchor, vocoded copies,
    augmented vocoded copies. The assembly reads originals and augmented waveforms in place through a row table (no regrouping
    copy). Returns ``(out, out_len)`` as :func:`assemble` (plus the [G, V] label tensor with ``with_labels``)."""
    drawn, starts, all_waves = [], [], []
    nvoc = len(items[0][1])
    for anchor, vocoded in items:
        if len(vocoded) != nvoc:
            raise ValueError("every item needs the same number of vocoded copies")
        order = list(vocoded) + [anchor]
        drawn += [_plans.draw_for_algo(w.shape[0], sr, args, 5) for w in order]
        all_waves += order
        starts.append(crop_plan(anchor.shape[0], trim_length, random_trim_nosil, repeat_pad)[0])
    bp = _plans.pack(drawn)
    x, ln = eng.pack_waveforms(all_waves, ld=bp.ld)
    y = eng.process(5, x, ln, eng.upload_plan(bp))
    G, V = len(items), 2 * (nvoc + 1)
    rows = torch.from_numpy(item_view_rows(G, nvoc).reshape(-1)).to(eng.device)
    label = torch.from_numpy(item_labels(nvoc)) if with_labels else None
    out, out_len, labels = assemble_ex(eng, x, y, rows, ln, V, starts, trim_length, repeat_pad, layout, view_label=label)
    return (out, out_len, labels) if with_labels else (out, out_len)


class ItemBatcher:
    """Loader-level drop-in for ``Dataset_for.__getitem__`` (asvspoof_2019_augall_3.py:103-146) over a BATCH of indices.

    The reference builds one item per call inside a DataLoader worker: RawBoost on every vocoded copy and on the anchor, a
    ``np.random.choice`` of additional bona fide utterances, one shared crop, a column concat. This class does the same for a
    list of indices with ONE device pass: every random draw is made on the host in the reference's order on the process-global
    numpy stream (so the stream ends where the reference leaves it), all RawBoost calls of all items run as one batch, and
    the views are assembled on the device in place. ``augmentation_methods`` is RawBoost12 only -- the other methods of the
    reference (background noise, reverb wrappers) need external corpora and stay with the caller.

    ``load_audio(path) -> float32 waveform`` is the caller's reader (``librosa.load`` in the reference)."""

    def __init__(self, args, list_ids, base_dir, load_audio, vocoders=(), num_additional_real=2, trim_length=64000, wav_samp_rate=16000,
                 repeat_pad=True, engine: Optional[Engine] = None, planner: str = "numpy"):
        import os
        self.args, self.list_ids, self.load_audio = args, list(list_ids), load_audio
        self.bonafide_dir = os.path.join(base_dir, "bonafide")
        self.vocoded_dir = os.path.join(base_dir, "vocoded")
        self.vocoders, self.num_additional_real = list(vocoders), int(num_additional_real)
        self.trim_length, self.sr, self.repeat_pad = int(trim_length), int(wav_samp_rate), bool(repeat_pad)
        self.eng = engine
        self.planner = planner
        self._native = None

    def _draw(self, length):
        if self.planner == "native":
            if self._native is None:
                from .native_planner import NativePlanner
                self._native = NativePlanner(threads=1, pinned=False)
            return self._native.draw([length], self.sr, self.args, 5, use_global_stream=True, copy=True)
        return _plans.draw_for_algo(length, self.sr, self.args, 5)

    def items(self, indices: Sequence[int], layout: int = LAYOUT_ITEM):
        """``[(utt_id, data, label)]``-equivalent for ``indices``: returns ``(ids, data, labels, out_len)`` with ``data`` a device
        tensor [G, trim_length, V] (``layout`` LAYOUT_ITEM, what ``Tensor(batch_data)`` holds per item) or [G, V, trim_length]."""
        import os
        eng = self.eng or default_engine()
        nvoc, n_add = len(self.vocoders), self.num_additional_real
        plans, waves, extra_waves, starts = [], [], [], []
        for idx in indices:
            name = self.list_ids[idx]
            anchor = np.asarray(self.load_audio(os.path.join(self.bonafide_dir, name)), dtype=np.float32)
            for v in self.vocoders:  # vocoded copies first: that is the order in which the loader consumes the stream
                w = np.asarray(self.load_audio(os.path.join(self.vocoded_dir, v + "_" + name)), dtype=np.float32)
                waves.append(w)
                plans.append(self._draw(w.shape[0]))
            waves.append(anchor)
            plans.append(self._draw(anchor.shape[0]))
            others = list(range(len(self.list_ids)))
            others.remove(idx)
            picked = np.random.choice(others, n_add, replace=False)
            extra_waves += [np.asarray(self.load_audio(os.path.join(self.bonafide_dir, self.list_ids[i])), dtype=np.float32) for i in picked]
            starts.append(crop_plan(anchor.shape[0], self.trim_length, True, self.repeat_pad)[0])
        G, per = len(indices), nvoc + 1
        if plans and isinstance(plans[0], _plans.BatchPlan):
            bp = _concat_batch_plans(plans)
        else:
            bp = _plans.pack(plans)
        # rows: the RawBoost inputs of all items first, then every item's additional bona fide utterances (not augmented)
        ld = _plans.padded_ld(max([w.shape[0] for w in waves + extra_waves] + [1]))
        bp = _repad(bp, ld)
        x, ln = eng.pack_waveforms(waves + extra_waves, ld=ld)
        y = torch.zeros_like(x)
        nb = G * per
        eng.process(5, x[:nb], ln[:nb], eng.upload_plan(bp), out=y[:nb])
        extra_rows = nb + np.arange(G * n_add).reshape(G, n_add) if n_add else None
        V = 2 + n_add + 2 * nvoc
        rows = torch.from_numpy(item_view_rows(G, nvoc, extra_rows).reshape(-1)).to(eng.device)
        label = torch.from_numpy(item_labels(nvoc, n_add))
        out, out_len, labels = assemble_ex(eng, x, y, rows, ln, V, starts, self.trim_length, self.repeat_pad, layout, view_label=label)
        return [self.list_ids[i] for i in indices], out, labels, out_len


def _repad(bp, ld):
    """A plan whose rows are addressed with stride ``ld`` (only SSI noise depends on the stride; algo 5 has none)."""
    bp.ld = int(ld)
    return bp


def _concat_batch_plans(parts):
    """Concatenate one-utterance BatchPlans (native planner, global stream) into one batch plan (LnL + ISD fields)."""
    from .plans import BatchPlan
    B = len(parts)
    out = BatchPlan(B=B, ld=max(p.ld for p in parts), lengths=np.concatenate([p.lengths for p in parts]).astype(np.int32),
                    g_sd=parts[0].g_sd, n_f=parts[0].n_f)
    out.lnl_taps = np.concatenate([p.lnl_taps for p in parts])
    sizes = np.concatenate([np.diff(p.lnl_tap_off) for p in parts])
    out.lnl_tap_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    out.isd_idx = np.concatenate([p.isd_idx for p in parts])
    out.isd_fr = np.concatenate([p.isd_fr for p in parts])
    out.isd_off = np.concatenate([[0], np.cumsum([p.isd_idx.shape[0] for p in parts])]).astype(np.int32)
    return out
