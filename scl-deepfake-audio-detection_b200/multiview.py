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
from typing import List, Optional, Sequence, Tuple

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


def item_views(eng: Engine, items: Sequence[Tuple[np.ndarray, Sequence[np.ndarray]]], args, sr: int, trim_length: int,
               repeat_pad: bool = True, random_trim_nosil: bool = True, layout: int = LAYOUT_MODEL):
    """Views of whole items, RawBoost and assembly on the device, random draws in the loader's order.

    ``items``: ``(anchor, [vocoded copies])`` float32 waveforms. Per item, as Dataset_for.__getitem__ with
    ``augmentation_methods[0] == 'RawBoost12'`` does (asvspoof_2019_augall_3.py:109-138): RawBoost (algo 5) is drawn for
    each vocoded copy, then for the anchor, then the shared crop. View order: anchor, augmented anchor, vocoded copies,
    augmented vocoded copies. Returns ``(out, out_len)`` as :func:`assemble`."""
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
    # rows per item in x / y: vocoded..., anchor. Views: anchor, aug anchor, vocoded..., aug vocoded...
    per, V = nvoc + 1, 2 * (nvoc + 1)
    G = len(items)
    xs, ys = x.view(G, per, -1), y.view(G, per, -1)
    views = torch.cat([xs[:, nvoc:], ys[:, nvoc:], xs[:, :nvoc], ys[:, :nvoc]], dim=1).reshape(G * V, -1).contiguous()
    l2 = ln.view(G, per)
    vlen = torch.cat([l2[:, nvoc:], l2[:, nvoc:], l2[:, :nvoc], l2[:, :nvoc]], dim=1).reshape(-1).contiguous()
    return assemble(eng, views, vlen, V, starts, trim_length, repeat_pad, layout)
