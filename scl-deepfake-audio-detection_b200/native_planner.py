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


_direct_ok = None


def _direct_exchange_ok() -> bool:
    """One-time self-check of the assumption the direct state exchange rests on: that ``bitgen.ctypes.state_address`` points at
    numpy's ``mt19937_state {uint32 key[624]; int pos;}``. A private generator is seeded, its state read through the public
    ``.state`` property and through the raw address, and the two must agree -- before and after drawing. On any mismatch (a
    numpy whose private layout changed) the slower public ``get_state`` / ``set_state`` path is used instead, for good."""
    global _direct_ok
    if _direct_ok is None:
        try:
            bg = np.random.MT19937(987654321)
            ok = True
            for _ in range(2):
                st = bg.state["state"]
                raw = (C.c_uint32 * 625).from_address(bg.ctypes.state_address)
                ok = ok and np.array_equal(np.frombuffer(raw, dtype=np.uint32, count=624), st["key"]) and int(np.int32(raw[624])) == int(st["pos"])
                bg.random_raw(700)  # crosses a block boundary: key and pos both change
            _direct_ok = bool(ok)
        except Exception:
            _direct_ok = False
    return _direct_ok


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
            direct = algo not in ALGO_USES_SSI and type(bitgen).__name__ == "MT19937" and _direct_exchange_ok()
            if direct:
                bitgen.lock.acquire()
                try:
                    state_addr = bitgen.ctypes.state_address
                    C.memmove(C.byref(state), state_addr, _MT_STATE_BYTES)
                    state.has_gauss, state.cached_gaussian = 0, 0.0
                except BaseException:
                    bitgen.lock.release()  # never leave numpy's global generator locked
                    raise
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

    def self_check(self) -> None:
        """Draw one short algo-4 utterance (LnL, ISD and SSI: every kind of draw) here and with the numpy calls themselves from
        the same seed and demand identical integers, impulse gains, float32 taps / noise and the same stream state afterwards.
        Raises :class:`RawBoostLibraryError` on any difference -- a numpy whose legacy generator changed must not silently
        desynchronise everything drawn after a RawBoost call. The caller's global stream is left untouched."""
        from types import SimpleNamespace
        from . import plans as _plans
        args = SimpleNamespace(N_f=2, nBands=2, minF=20, maxF=8000, minBW=100, maxBW=1000, minCoeff=10, maxCoeff=40, minG=0, maxG=0,
                               minBiasLinNonLin=5, maxBiasLinNonLin=20, P=10, g_sd=2, SNRmin=10, SNRmax=40)
        saved = np.random.get_state()
        try:
            problems = []
            for algo in (4, 5):  # 5 goes through the direct state exchange, 4 (SSI) through get_state / set_state
                np.random.seed(20240229)
                want = _plans.pack([_plans.draw_for_algo(1500, 16000, args, algo)])
                want_state = np.random.get_state()
                np.random.seed(20240229)
                got = self.draw([1500], 16000, args, algo, use_global_stream=True, copy=True)
                got_state = np.random.get_state()
                for name in ("lnl_tap_off", "isd_off", "isd_idx", "isd_fr", "ssi_tap_off", "ssi_snr_db"):
                    w, g = getattr(want, name), getattr(got, name)
                    if (w is None) != (g is None) or (w is not None and not np.array_equal(w, g)):
                        problems.append(f"algo {algo}: {name}")
                for name in ("lnl_taps", "ssi_taps", "ssi_noise"):
                    w, g = getattr(want, name), getattr(got, name)
                    if (w is None) != (g is None):
                        problems.append(f"algo {algo}: {name}")
                    elif w is not None and (w.shape != g.shape or not np.all(np.abs(w.astype(np.float64) - g) <= np.spacing(np.abs(w)).astype(np.float64))):
                        problems.append(f"algo {algo}: {name}")
                if not (np.array_equal(want_state[1], got_state[1]) and want_state[2:] == got_state[2:]):
                    problems.append(f"algo {algo}: stream state after the call")
        finally:
            np.random.set_state(saved)
        if problems:
            raise _lib.RawBoostLibraryError(
                f"native planner does not reproduce numpy {np.__version__}'s legacy random stream ({', '.join(problems)}); "
                "set RAWBOOST_B200_PLANNER=numpy to draw with the numpy calls themselves")

    def close(self):
        if self._h is not None:
            self.lib.rb_planner_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
