"""Native (C++) plan drawing: the fast equivalent of :func:`plans.draw_batch`.

``lib/librawboost_b200.so`` re-implements, bit for bit, the numpy legacy MT19937 calls the reference makes
(``/root/reference/datautils/RawBoost.py:15,79,80,90``) and the float64 filter design of ``genNotchCoeffs``
(RawBoost.py:28-48); see ``csrc/rb_planner.cpp``. Tap counts, impulse counts / positions / gains, SSI noise and the stream
state are identical to numpy's; tap values agree to ~1e-15 relative before the float32 cast. The numpy path in
:mod:`plans` remains the contract; this is the throughput path (host threads, page-locked output buffers, no GIL).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .plans import ALGO_USES_ISD, ALGO_USES_LNL, ALGO_USES_SSI, BatchPlan, padded_ld


_args_struct = _lib.args_struct


_MT_STATE_BYTES = 624 * 4 + 4  # numpy's mt19937_state: uint32 key[624]; int pos (numpy/random/src/mt19937/mt19937.h)


def _view(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
    return np.frombuffer((C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr), dtype=dtype)  # a view, no copy


class NativePlanner:
    """Owns one set of (page-locked when a GPU is present) plan buffers; each :meth:`draw` overwrites them."""

    def __init__(self, threads: int = 0, pinned: bool = True):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.rb_planner_create(C.byref(h), int(threads), int(bool(pinned))), "rb_planner_create")
        self._h = h

    def draw(self, lengths: Sequence[int], sr, args, algo: int, seeds: Optional[Sequence[int]] = None, ld: Optional[int] = None,
             use_global_stream: bool = False, copy: bool = False) -> BatchPlan:
        """Plans of a batch. ``seeds``: re-seed before each utterance (parallel). ``use_global_stream``: consume numpy's
        process-global stream sequentially and leave it exactly where the reference's calls would (``seeds`` must be None)."""
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        B = int(lengths.shape[0])
        ld = int(ld or padded_ld(int(lengths.max()) if B else 0))
        a = _args_struct(args, sr)
        view = _lib.RbPlan()
        state = None
        seeds_arr = None
        if seeds is not None:
            seeds_arr = np.ascontiguousarray(seeds, dtype=np.uint32)
            if seeds_arr.shape[0] != B:
                raise ValueError("one seed per utterance")
        elif use_global_stream:
            state = _lib.RbRngState()
            bitgen = np.random.mtrand._rand._bit_generator
            # Without SSI no normal is drawn, so the legacy Gaussian cache is neither read nor written and the 624 words + cursor
            # can be exchanged directly with the bit generator's own mt19937_state {uint32 key[624]; int pos;} under its
            # lock (np.random.get_state() / set_state() cost ~70 us each, more than the LnL draw itself).
            direct = algo not in ALGO_USES_SSI and type(bitgen).__name__ == "MT19937"
            if direct:
                bitgen.lock.acquire()
                state_addr = bitgen.ctypes.state_address
                C.memmove(C.byref(state), state_addr, _MT_STATE_BYTES)
                state.has_gauss, state.cached_gaussian = 0, 0.0
            else:
                name, key, pos, has_gauss, cached = np.random.get_state()
                C.memmove(state.key, np.ascontiguousarray(key, dtype=np.uint32).ctypes.data, 624 * 4)
                state.pos, state.has_gauss, state.cached_gaussian = int(pos), int(has_gauss), float(cached)
        else:
            raise ValueError("give per-utterance seeds or use_global_stream=True")
        try:
            rc = self.lib.rb_planner_draw(self._h, C.byref(a), int(algo), B, ld, C.c_void_p(lengths.ctypes.data),
                                          C.c_void_p(seeds_arr.ctypes.data) if seeds_arr is not None else None,
                                          C.byref(state) if state is not None else None, C.byref(view))
            _lib.check(rc, "rb_planner_draw")
            if state is not None and direct:
                C.memmove(state_addr, C.byref(state), _MT_STATE_BYTES)
        finally:
            if state is not None and direct:
                bitgen.lock.release()
        if state is not None and not direct:
            key = np.frombuffer(bytes(state.key), dtype=np.uint32).copy()
            np.random.set_state(("MT19937", key, int(state.pos), int(state.has_gauss), float(state.cached_gaussian)))
        bp = BatchPlan(B=B, ld=ld, lengths=lengths, g_sd=float(args.g_sd) if algo in ALGO_USES_ISD else 0.0)
        fin = (lambda arr: arr.copy()) if copy else (lambda arr: arr)
        if algo in ALGO_USES_LNL and B:
            bp.n_f = int(args.N_f)
            bp.lnl_tap_off = fin(_view(view.lnl_tap_off, B * bp.n_f + 1, np.int32))
            bp.lnl_taps = fin(_view(view.lnl_taps, int(bp.lnl_tap_off[-1]), np.float32))
        if algo in ALGO_USES_ISD and B:
            bp.isd_off = fin(_view(view.isd_off, B + 1, np.int32))
            n = int(bp.isd_off[-1])
            bp.isd_idx = fin(_view(view.isd_idx, n, np.int32))
            bp.isd_fr = fin(_view(view.isd_fr, n, np.float64))
        if algo in ALGO_USES_SSI and B:
            bp.ssi_noise = fin(_view(view.ssi_noise, B * ld, np.float32).reshape(B, ld))
            bp.ssi_tap_off = fin(_view(view.ssi_tap_off, B + 1, np.int32))
            bp.ssi_taps = fin(_view(view.ssi_taps, int(bp.ssi_tap_off[-1]), np.float32))
            bp.ssi_snr_db = fin(_view(view.ssi_snr_db, B, np.float32))
        return bp

    def close(self):
        if self._h is not None:
            self.lib.rb_planner_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
