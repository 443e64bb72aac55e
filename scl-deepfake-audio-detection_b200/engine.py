"""Batched device execution of RawBoost plans through the C ABI.

PyTorch is plumbing only here (device memory, streams, pinned buffers); all arithmetic happens in
``lib/librawboost_b200.so``. There is no CPU fallback: constructing an :class:`Engine` without CUDA raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .plans import BatchPlan, padded_ld


@dataclass
class DevicePlan:
    """A :class:`BatchPlan` whose arrays live on the device; ``struct`` is the ``rb_plan`` passed to the ABI."""
    B: int
    ld: int
    lengths: torch.Tensor
    tensors: dict
    struct: _lib.RbPlan
    host: Optional[BatchPlan] = None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """One engine per (process, device). Owns a growable device workspace; thread-compatible, not thread-safe."""

    def __init__(self, device: int | str | torch.device = 0):
        if not torch.cuda.is_available():
            raise _lib.RawBoostLibraryError("rawboost_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        major, _ = torch.cuda.get_device_capability(self.device)
        if major != 10:
            raise _lib.RawBoostLibraryError(f"rawboost_b200 is built for sm_100a only; device capability is {major}.x")
        self._ws: Optional[torch.Tensor] = None
        self._ctx = None

    # -- memory ---------------------------------------------------------------------------------------
    def workspace(self, B: int, ld: int, algo: int = 8) -> torch.Tensor:
        """Growable device workspace, sized for what ``algo`` needs (algo 8 needs the most)."""
        need = int(self.lib.rb_workspace_bytes_for(int(algo), B, ld))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _ws_ptr(ws: torch.Tensor) -> int:
        return (ws.data_ptr() + 255) // 256 * 256

    def upload_plan(self, plan: BatchPlan, non_blocking: bool = True) -> DevicePlan:
        """Copy the CSR arrays of a host plan to the device (pinned staging when available)."""
        t = {}
        for name in ("lnl_taps", "lnl_tap_off", "isd_off", "isd_idx", "isd_fr", "ssi_noise", "ssi_taps", "ssi_tap_off", "ssi_snr_db"):
            arr = getattr(plan, name)
            if arr is not None:
                src = torch.from_numpy(np.ascontiguousarray(arr))
                t[name] = src.to(self.device, non_blocking=False)
            else:
                t[name] = None
        lengths = torch.from_numpy(np.ascontiguousarray(plan.lengths)).to(self.device)
        s = _lib.RbPlan()
        s.n_f = int(plan.n_f)
        s.g_sd = float(plan.g_sd)
        for name, v in t.items():
            setattr(s, name, None if v is None else v.data_ptr())
        return DevicePlan(B=plan.B, ld=plan.ld, lengths=lengths, tensors=t, struct=s, host=plan)

    def draw_device_plan(self, lengths: torch.Tensor, seeds, sr, args, algo: int, ld: int) -> DevicePlan:
        """Draw the plans of a seeded batch ON THE DEVICE (``rb_devplan_draw``): ``np.random.seed(seeds[u])`` precedes
        utterance u, exactly as ``plans.draw_batch(..., seeds=...)`` / ``NativePlanner.draw`` do on the host. Asynchronous on
        torch's current stream; the returned plan's arrays live in one device buffer owned by the plan."""
        B = int(lengths.numel())
        if not torch.is_tensor(seeds):
            seeds = torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.uint32).view(np.int32)).to(self.device)
        a = _lib.args_struct(args, sr)
        need = int(self.lib.rb_devplan_bytes(C.byref(a), int(algo), B, int(ld)))
        store = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        s = _lib.RbPlan()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.rb_devplan_draw(C.byref(a), int(algo), B, int(ld), _ptr(lengths), _ptr(seeds), C.c_void_p(self._ws_ptr(store)),
                                          need, C.byref(s), C.c_void_p(stream))
        _lib.check(rc, f"rb_devplan_draw(algo={algo})")
        return DevicePlan(B=B, ld=int(ld), lengths=lengths, tensors={"storage": store, "seeds": seeds}, struct=s, host=None)

    def download_plan(self, dp: DevicePlan) -> BatchPlan:
        """Copy a device-drawn plan back to the host (tests / inspection)."""
        store = dp.tensors["storage"]

        def fetch(ptr, count, dtype):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dtype)
            off = int(ptr) - store.data_ptr()
            return store[off:off + count * np.dtype(dtype).itemsize].cpu().numpy().view(dtype).copy()

        s, B, ld = dp.struct, dp.B, dp.ld
        bp = BatchPlan(B=B, ld=ld, lengths=dp.lengths.cpu().numpy(), g_sd=float(s.g_sd))
        if s.lnl_tap_off:
            bp.n_f = int(s.n_f)
            bp.lnl_tap_off = fetch(s.lnl_tap_off, B * bp.n_f + 1, np.int32)
            bp.lnl_taps = fetch(s.lnl_taps, int(bp.lnl_tap_off[-1]), np.float32)
        if s.isd_off:
            bp.isd_off = fetch(s.isd_off, B + 1, np.int32)
            bp.isd_idx = fetch(s.isd_idx, int(bp.isd_off[-1]), np.int32)
            bp.isd_fr = fetch(s.isd_fr, int(bp.isd_off[-1]), np.float64)
        if s.ssi_tap_off:
            bp.ssi_noise = fetch(s.ssi_noise, B * ld, np.float32).reshape(B, ld)
            bp.ssi_tap_off = fetch(s.ssi_tap_off, B + 1, np.int32)
            bp.ssi_taps = fetch(s.ssi_taps, int(bp.ssi_tap_off[-1]), np.float32)
            bp.ssi_snr_db = fetch(s.ssi_snr_db, B, np.float32)
        return bp

    def pack_waveforms(self, waves: Sequence[np.ndarray], ld: Optional[int] = None):
        """Host list of 1-D float arrays -> ([B, ld] float32 device tensor, int32 lengths on device)."""
        lengths = np.array([int(w.shape[0]) for w in waves], dtype=np.int32)
        ld = ld or padded_ld(int(lengths.max()) if len(waves) else 0)
        host = np.zeros((len(waves), ld), dtype=np.float32)
        for u, w in enumerate(waves):
            host[u, :lengths[u]] = np.asarray(w, dtype=np.float32)
        return torch.from_numpy(host).to(self.device), torch.from_numpy(lengths).to(self.device)

    # -- execution ------------------------------------------------------------------------------------
    def process(self, algo: int, x: torch.Tensor, lengths: torch.Tensor, plan: Optional[DevicePlan],
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``process_Rawboost_feature`` for a whole batch on the device: x [B, ld] float32 -> new [B, ld] tensor.

        Asynchronous on torch's current stream. Samples beyond ``lengths[u]`` of the output row are zero when the
        tensor is allocated here, untouched when ``out`` is passed."""
        self._check_batch(x, lengths)
        B, ld = x.shape
        if plan is not None and (plan.B != B or plan.ld != ld):  # CSR offsets and the SSI noise rows are laid out for (B, ld)
            raise ValueError(f"plan was packed for B={plan.B}, ld={plan.ld}; the batch is B={B}, ld={ld}")
        if out is None:
            out = torch.zeros_like(x)
        ws = self.workspace(B, ld, algo)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ps = C.byref(plan.struct) if plan is not None else None
        with torch.cuda.device(self.device):
            rc = self.lib.rb_process(int(algo), _ptr(x), _ptr(lengths), B, ld, ps, _ptr(out), C.c_void_p(self._ws_ptr(ws)),
                                     ws.numel() - 256, C.c_void_p(stream))
        _lib.check(rc, f"rb_process(algo={algo})")
        return out

    def filter_fir(self, x: torch.Tensor, lengths: torch.Tensor, taps: torch.Tensor, tap_off: torch.Tensor,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._check_batch(x, lengths)
        B, ld = x.shape
        if out is None:
            out = torch.zeros_like(x)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.rb_filter_fir(_ptr(x), _ptr(lengths), B, ld, _ptr(taps), _ptr(tap_off), _ptr(out), C.c_void_p(stream))
        _lib.check(rc, "rb_filter_fir")
        return out

    def normwav(self, x: torch.Tensor, lengths: torch.Tensor, always: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._check_batch(x, lengths)
        B, ld = x.shape
        if out is None:
            out = torch.zeros_like(x)
        ws = self.workspace(B, ld, 0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.rb_normwav(_ptr(x), _ptr(lengths), B, ld, int(bool(always)), _ptr(out), C.c_void_p(self._ws_ptr(ws)),
                                     ws.numel() - 256, C.c_void_p(stream))
        _lib.check(rc, "rb_normwav")
        return out

    def _check_batch(self, x: torch.Tensor, lengths: torch.Tensor) -> None:
        if x.device != self.device or lengths.device != self.device:
            raise ValueError(f"tensors must live on {self.device}")
        if x.dtype != torch.float32 or x.dim() != 2 or not x.is_contiguous():
            raise ValueError("x must be a contiguous [B, ld] float32 tensor")
        if lengths.dtype != torch.int32 or lengths.numel() != x.shape[0]:
            raise ValueError("lengths must be int32 [B]")

    # -- host-buffer path (rb_process_host): what a non-torch caller of the C ABI would use ---------------
    def process_host(self, algo: int, x: np.ndarray, plan: Optional[BatchPlan], out: Optional[np.ndarray] = None) -> np.ndarray:
        """x: [B, ld] float32 host array (pinned for asynchronous DMA); plan: host :class:`BatchPlan`.
        Copies in, computes, copies out; returns when ``out`` is complete."""
        self._host_ctx()
        if x.dtype != np.float32 or x.ndim != 2 or not x.flags.c_contiguous:
            raise ValueError("x must be a C-contiguous [B, ld] float32 array")
        B, ld = x.shape
        if plan is not None and (plan.B != B or plan.ld != ld):
            raise ValueError(f"plan was packed for B={plan.B}, ld={plan.ld}; the batch is B={B}, ld={ld}")
        if out is None:
            out = np.zeros_like(x)
        s = _lib.RbPlan()
        keep = []
        lengths = plan.lengths if plan is not None else np.full(B, ld, dtype=np.int32)
        if plan is not None:
            s.n_f = int(plan.n_f)
            s.g_sd = float(plan.g_sd)
            for name in ("lnl_taps", "lnl_tap_off", "isd_off", "isd_idx", "isd_fr", "ssi_noise", "ssi_taps", "ssi_tap_off", "ssi_snr_db"):
                arr = getattr(plan, name)
                if arr is not None:
                    arr = np.ascontiguousarray(arr)
                    keep.append(arr)
                    setattr(s, name, arr.ctypes.data)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        rc = self.lib.rb_process_host(self._ctx, int(algo), C.c_void_p(x.ctypes.data), C.c_void_p(lengths.ctypes.data), B, ld,
                                      C.byref(s), C.c_void_p(out.ctypes.data))
        _lib.check(rc, f"rb_process_host(algo={algo})")
        return out

    def _host_ctx(self):
        if self._ctx is None:
            ctx = C.c_void_p()
            _lib.check(self.lib.rb_ctx_create(C.byref(ctx), self.device.index or 0), "rb_ctx_create")
            self._ctx = ctx
        return self._ctx

    def set_host_chunk(self, utterances: int) -> None:
        """Utterances per pipeline chunk of the host-buffer entry points (0 = default: four per SM)."""
        _lib.check(self.lib.rb_ctx_set_chunk(self._host_ctx(), int(utterances)), "rb_ctx_set_chunk")

    def set_host_plan_mode(self, mode: int) -> None:
        """Device planner placement in :meth:`process_host_seeded`: 0 = on its own streams beside the kernels (default), 1 = in line."""
        _lib.check(self.lib.rb_ctx_set_plan_mode(self._host_ctx(), int(mode)), "rb_ctx_set_plan_mode")

    def process_host_seeded(self, algo: int, x: np.ndarray, lengths: np.ndarray, seeds, sr, args,
                            out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host waveforms in, host results out, plans drawn ON THE DEVICE from per-utterance seeds
        (``np.random.seed(seeds[u])`` before utterance u). x: [B, ld] float32, page-locked for full overlap."""
        if x.dtype != np.float32 or x.ndim != 2 or not x.flags.c_contiguous:
            raise ValueError("x must be a C-contiguous [B, ld] float32 array")
        B, ld = x.shape
        if out is None:
            out = np.zeros_like(x)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        if lengths.shape[0] != B or seeds.shape[0] != B:
            raise ValueError("one length and one seed per utterance")
        a = _lib.args_struct(args, sr)
        rc = self.lib.rb_process_host_seeded(self._host_ctx(), int(algo), C.byref(a), C.c_void_p(x.ctypes.data),
                                             C.c_void_p(lengths.ctypes.data), C.c_void_p(seeds.ctypes.data), B, ld,
                                             C.c_void_p(out.ctypes.data))
        _lib.check(rc, f"rb_process_host_seeded(algo={algo})")
        return out

    def submit_host_seeded(self, algo: int, x: np.ndarray, lengths: np.ndarray, seeds: np.ndarray, sr, args, out: np.ndarray) -> int:
        """Streaming form of :meth:`process_host_seeded`: queue the call and return a ticket at once. Up to two calls may be in
        flight; ``x``, ``lengths``, ``seeds`` and ``out`` (contiguous, ideally page-locked, of the exact dtypes int32 / uint32 /
        float32) must stay alive and untouched until :meth:`wait_host` has returned for the ticket."""
        if x.dtype != np.float32 or x.ndim != 2 or not x.flags.c_contiguous or out.dtype != np.float32 or out.shape != x.shape:
            raise ValueError("x / out must be C-contiguous [B, ld] float32 arrays of the same shape")
        if lengths.dtype != np.int32 or seeds.dtype != np.uint32 or not lengths.flags.c_contiguous or not seeds.flags.c_contiguous:
            raise ValueError("lengths must be a contiguous int32 array and seeds a contiguous uint32 array")
        B, ld = x.shape
        a = _lib.args_struct(args, sr)
        ticket = C.c_uint64(0)
        rc = self.lib.rb_submit_host_seeded(self._host_ctx(), int(algo), C.byref(a), C.c_void_p(x.ctypes.data),
                                            C.c_void_p(lengths.ctypes.data), C.c_void_p(seeds.ctypes.data), B, ld,
                                            C.c_void_p(out.ctypes.data), C.byref(ticket))
        _lib.check(rc, f"rb_submit_host_seeded(algo={algo})")
        return int(ticket.value)

    def submit_host_ex(self, algo: int, x: np.ndarray, dtype: str, lengths: np.ndarray, seeds: np.ndarray, sr, args, out) -> int:
        """General streaming submit (``rb_submit_seeded_ex``) for host input: ``dtype`` "f32" (float32 waveforms) or "pcm16"
        (int16 samples as a wav file holds them; converted on the device as sample / 32768, which is what ``librosa.load``
        returns for 16-bit audio). ``out``: a host float32 array [B, ld] or a device tensor [B, ld] -- then the results stay on
        the device and nothing is copied back. Returns a ticket for :meth:`wait_host`; buffers must stay alive until then."""
        want = np.int16 if dtype == "pcm16" else np.float32
        if x.dtype != want or x.ndim != 2 or not x.flags.c_contiguous:
            raise ValueError(f"x must be a C-contiguous [B, ld] {np.dtype(want).name} array")
        if lengths.dtype != np.int32 or seeds.dtype != np.uint32 or not lengths.flags.c_contiguous or not seeds.flags.c_contiguous:
            raise ValueError("lengths must be a contiguous int32 array and seeds a contiguous uint32 array")
        B, ld = x.shape
        if torch.is_tensor(out):
            if out.device != self.device or out.dtype != torch.float32 or tuple(out.shape) != (B, ld) or not out.is_contiguous():
                raise ValueError("a device sink must be a contiguous [B, ld] float32 tensor on the engine's device")
            y_ptr, y_kind = out.data_ptr(), _lib.RB_IO_DEVICE_F32
        else:
            if out.dtype != np.float32 or out.shape != (B, ld) or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous [B, ld] float32 array")
            y_ptr, y_kind = out.ctypes.data, _lib.RB_IO_HOST_F32
        a = _lib.args_struct(args, sr)
        ticket = C.c_uint64(0)
        rc = self.lib.rb_submit_seeded_ex(self._host_ctx(), int(algo), C.byref(a), C.c_void_p(x.ctypes.data),
                                          _lib.RB_IO_HOST_PCM16 if dtype == "pcm16" else _lib.RB_IO_HOST_F32,
                                          C.c_void_p(lengths.ctypes.data), C.c_void_p(seeds.ctypes.data), B, ld, C.c_void_p(y_ptr), y_kind,
                                          None, 0, C.byref(ticket))
        _lib.check(rc, f"rb_submit_seeded_ex(algo={algo})")
        return int(ticket.value)

    def process_device_seeded(self, algo: int, x: torch.Tensor, lengths: torch.Tensor, seeds: torch.Tensor, sr, args,
                              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Seeded batch that already lives on the device: plans drawn on the device chunk by chunk WHILE the previous chunk is
        being filtered (the planner's latency-bound kernels run beside the FIR kernel on their own low-priority streams),
        results left on the device. Ordered after the work queued on torch's current stream, which in turn waits for the
        results: no host synchronisation. ``seeds``: int32 / uint32 device tensor (``np.random.seed(seeds[u])`` per utterance)."""
        self._check_batch(x, lengths)
        B, ld = x.shape
        if out is None:
            out = torch.zeros_like(x)
        if seeds.device != self.device or seeds.numel() != B or seeds.element_size() != 4:
            raise ValueError("seeds must be a 32-bit integer tensor [B] on the engine's device")
        a = _lib.args_struct(args, sr)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ticket = C.c_uint64(0)
        with torch.cuda.device(self.device):
            rc = self.lib.rb_submit_seeded_ex(self._host_ctx(), int(algo), C.byref(a), C.c_void_p(x.data_ptr()), _lib.RB_IO_DEVICE_F32,
                                              C.c_void_p(lengths.data_ptr()), C.c_void_p(seeds.data_ptr()), B, ld,
                                              C.c_void_p(out.data_ptr()), _lib.RB_IO_DEVICE_F32, C.c_void_p(stream), 1, C.byref(ticket))
        _lib.check(rc, f"rb_submit_seeded_ex(device, algo={algo})")
        return out

    def wait_host(self, ticket: int = 0) -> None:
        """Block until the submitted call ``ticket`` (0: every submitted call) has delivered its results."""
        _lib.check(self.lib.rb_ctx_wait(self._host_ctx(), int(ticket)), "rb_ctx_wait")

    def trace_host(self, on: bool) -> None:
        """Record a per-chunk timeline of the following host-buffer calls (see :meth:`host_timeline`)."""
        _lib.check(self.lib.rb_ctx_trace(self._host_ctx(), int(bool(on))), "rb_ctx_trace")

    def host_timeline(self):
        """Timeline of the last traced host-buffer call: one dict per pipeline chunk, times in ms from the call's start."""
        n = self.lib.rb_ctx_timeline(self._host_ctx(), None, 0)
        if n <= 0:
            return []
        buf = (C.c_double * n)()
        self.lib.rb_ctx_timeline(self._host_ctx(), buf, n)
        keys = ("first", "count", "copy_in_ms", "plan_ms", "kernels_ms", "copy_out_ms")
        return [dict(zip(keys, buf[i:i + 6])) for i in range(0, n, 6)]

    def last_host_traffic(self):
        h2d, d2h = C.c_uint64(0), C.c_uint64(0)
        if self._ctx is not None:
            _lib.check(self.lib.rb_ctx_last_traffic(self._ctx, C.byref(h2d), C.byref(d2h)))
        return int(h2d.value), int(d2h.value)

    def close(self):
        if self._ctx is not None:
            self.lib.rb_ctx_destroy(self._ctx)
            self._ctx = None
        self._ws = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_default: dict = {}


def default_device() -> int:
    """The device every per-utterance drop-in function uses: ``RAWBOOST_B200_DEVICE`` (default 0), read in ONE place."""
    import os
    return int(os.environ.get("RAWBOOST_B200_DEVICE", "0"))


def default_engine(device: Optional[int] = None) -> Engine:
    """Process-wide engine used by the per-utterance reference-style functions in :mod:`RawBoost`, :mod:`multiview` and
    :mod:`reverb` (``device`` None = :func:`default_device`)."""
    if device is None:
        device = default_device()
    eng = _default.get(device)
    if eng is None:
        eng = _default[device] = Engine(device)
    return eng
