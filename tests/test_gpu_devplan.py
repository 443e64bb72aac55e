"""GPU parity of the device-side planner (rb_devplan_draw) and of the pipelined host-buffer entry points.

The contract for random parameters is numpy's legacy global stream (SURVEY.md 8b "RNG convention"). The device planner
replays that stream on the GPU for independently seeded utterances; here it is held to:
  * integer work bit-exact against numpy itself (plans.draw_batch issues the reference's numpy calls): tap counts, impulse
    counts and positions; the float64 impulse gains bit-exact too (only exactly-rounded fp64 arithmetic is involved);
  * float32 taps / SSI noise within 1 ulp(fp32) of numpy's (sin/cos/log/pow may differ in the last fp64 bit);
  * results through the pipelined host entries bit-identical to the device-resident path on the same plans.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import rawboost_oracle as orc  # noqa: E402  (checker only)

ARGS = orc.make_args()
RAGGED = [1, 2, 3, 31, 32, 33, 34, 37, 63, 64, 65, 600, 1023, 1024, 1025, 2560, 2561, 4097, 16000, 40000, 49152, 49153, 65535, 65536]


@pytest.fixture(scope="module")
def eng():
    from scl_deepfake_audio_detection_b200.engine import Engine
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return Engine(0)


@pytest.fixture(scope="module")
def P():
    from scl_deepfake_audio_detection_b200 import plans
    return plans


def ulp_err(got, ref):
    got, ref = np.asarray(got, np.float32), np.asarray(ref, np.float32)
    if ref.size == 0:
        return 0.0
    return float(np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64)) / np.spacing(np.maximum(np.abs(ref), np.float32(1e-30)))))


def device_plan(eng, lengths, seeds, algo, args=ARGS):
    ld = (max(lengths) + 3) // 4 * 4
    ln = torch.tensor(lengths, dtype=torch.int32, device="cuda")
    dp = eng.draw_device_plan(ln, seeds, 16000, args, algo, ld)
    torch.cuda.synchronize()
    return dp, eng.download_plan(dp), ld


@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5, 6, 7, 8])
def test_device_plan_equals_numpy_plan(eng, P, algo):
    lengths = RAGGED if algo in (2, 5) else RAGGED[::3] + [16000]
    seeds = [(977 * algo + 13 * u) % 2 ** 32 for u in range(len(lengths))]
    _, got, ld = device_plan(eng, lengths, seeds, algo)
    ref = P.draw_batch(lengths, 16000, ARGS, algo, seeds=seeds, ld=ld)
    for name in ("lnl_tap_off", "isd_off", "isd_idx", "ssi_tap_off"):
        r = getattr(ref, name)
        if r is not None:
            assert np.array_equal(r, getattr(got, name)), f"{name} differs from numpy (integer work must be bit-exact)"
    if ref.isd_fr is not None:
        assert np.array_equal(ref.isd_fr, got.isd_fr), "impulse gains (float64) differ from numpy"
    if ref.ssi_snr_db is not None:
        assert np.array_equal(ref.ssi_snr_db, got.ssi_snr_db)
    for name in ("lnl_taps", "ssi_taps", "ssi_noise"):
        r = getattr(ref, name)
        if r is not None:
            assert ulp_err(getattr(got, name), r) <= 1.0, f"{name} more than 1 ulp(fp32) from numpy"


def test_device_plan_large_seeds_and_full_length(eng, P):
    """Seeds across the uint32 range; utterances of the benchmark's size."""
    lengths = [64600] * 6
    seeds = [0, 1, 1234, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1]
    _, got, ld = device_plan(eng, lengths, seeds, 5)
    ref = P.draw_batch(lengths, 16000, ARGS, 5, seeds=seeds, ld=ld)
    assert np.array_equal(ref.lnl_tap_off, got.lnl_tap_off)
    assert np.array_equal(ref.isd_off, got.isd_off)
    assert np.array_equal(ref.isd_idx, got.isd_idx)
    assert np.array_equal(ref.isd_fr, got.isd_fr)
    assert ulp_err(got.lnl_taps, ref.lnl_taps) <= 1.0


def test_device_plan_nondefault_arguments(eng, P):
    args = orc.make_args(N_f=3, nBands=6, minCoeff=20, maxCoeff=150, P=25, minG=-3, maxG=2, SNRmin=5, SNRmax=15)
    lengths = [5000, 64600, 777, 12345]
    seeds = [5, 6, 7, 8]
    for algo in (4, 5):
        _, got, ld = device_plan(eng, lengths, seeds, algo, args)
        ref = P.draw_batch(lengths, 16000, args, algo, seeds=seeds, ld=ld)
        assert np.array_equal(ref.lnl_tap_off, got.lnl_tap_off)
        assert np.array_equal(ref.isd_off, got.isd_off)
        assert np.array_equal(ref.isd_idx, got.isd_idx)
        assert np.array_equal(ref.isd_fr, got.isd_fr)
        assert ulp_err(got.lnl_taps, ref.lnl_taps) <= 1.0
        if algo == 4:
            assert np.array_equal(ref.ssi_tap_off, got.ssi_tap_off)
            assert ulp_err(got.ssi_taps, ref.ssi_taps) <= 1.0
            assert ulp_err(got.ssi_noise, ref.ssi_noise) <= 1.0


def test_device_plan_unsupported_is_loud(eng):
    """What the device planner does not implement is refused, not silently mis-drawn (cascades beyond freqz's 1024-point path)."""
    from scl_deepfake_audio_detection_b200 import _lib
    ln = torch.tensor([7000], dtype=torch.int32, device="cuda")
    with pytest.raises(_lib.RawBoostLibraryError):
        eng.draw_device_plan(ln, [1], 16000, orc.make_args(nBands=12, maxCoeff=200), 5, 7000)


@pytest.mark.parametrize("lengths", [[65537], [211000, 100000, 70001, 65537, 300, 64600], [49152, 49153, 50176, 50177, 65536]])
def test_device_plan_long_rows_equal_numpy(eng, P, lengths):
    """Un-cropped utterances (asvspoof_2019_augall_3.py:105-117 applies RawBoost before the crop): beyond 65536 samples the
    permutation runs as uint32 in global memory, beyond 49152 its first steps do -- impulse positions must stay numpy's, bit
    for bit, on both sides of both thresholds and in a ragged batch that mixes them."""
    seeds = [4000 + 7 * u for u in range(len(lengths))]
    for algo in (2, 5):
        _, got, ld = device_plan(eng, lengths, seeds, algo)
        ref = P.draw_batch(lengths, 16000, ARGS, algo, seeds=seeds, ld=ld)
        assert np.array_equal(ref.isd_off, got.isd_off)
        assert np.array_equal(ref.isd_idx, got.isd_idx), "impulse positions differ from numpy's permutation"
        assert np.array_equal(ref.isd_fr, got.isd_fr)
        if algo == 5:
            assert np.array_equal(ref.lnl_tap_off, got.lnl_tap_off) and ulp_err(got.lnl_taps, ref.lnl_taps) <= 1.0


def test_seeded_host_entry_long_rows(eng, P):
    """The pipelined host entry on a ragged batch of long rows == the resident path on numpy-drawn plans."""
    rs = np.random.RandomState(5)
    lengths = [120000, 64600, 211000, 65537, 9]
    waves = [(0.4 * rs.standard_normal(n)).astype(np.float32) for n in lengths]
    seeds = [900 + u for u in range(len(lengths))]
    for algo in (5, 2):
        bp = P.draw_batch(lengths, 16000, ARGS, algo, seeds=seeds)
        x, ln = eng.pack_waveforms(waves, ld=bp.ld)
        ref = eng.process(algo, x, ln, eng.upload_plan(bp)).cpu().numpy()
        got = eng.process_host_seeded(algo, x.cpu().numpy(), np.array(lengths, np.int32), np.array(seeds, np.uint32), 16000, ARGS)
        for u, n in enumerate(lengths):
            assert np.array_equal(got[u, :n], ref[u, :n]), (algo, u)


@pytest.mark.parametrize("algo", [0, 1, 2, 3, 5, 8])
@pytest.mark.parametrize("chunk", [1, 3, 0])
def test_seeded_host_entry_matches_device_path(eng, P, algo, chunk):
    """rb_process_host_seeded (device-drawn plans, chunked pipeline) == rb_process on numpy-drawn plans; and the
    pipelined rb_process_host (CSR plan sliced per chunk) == the same."""
    rs = np.random.RandomState(99 + algo)
    lengths = [4097, 64600, 1, 2561, 30000, 64600, 37, 12000]
    waves = [(0.5 * rs.standard_normal(n)).astype(np.float32) for n in lengths]
    seeds = [31 * algo + u for u in range(len(lengths))]
    bp = P.draw_batch(lengths, 16000, ARGS, algo if algo else 5, seeds=seeds)
    ld = bp.ld
    x, ln = eng.pack_waveforms(waves, ld=ld)
    ref = eng.process(algo, x, ln, eng.upload_plan(bp)).cpu().numpy()
    xh = x.cpu().numpy()
    eng.set_host_chunk(chunk)
    try:
        y_seeded = eng.process_host_seeded(algo, xh, np.array(lengths, np.int32), seeds, 16000, ARGS)
        y_plan = eng.process_host(algo, xh, bp)
    finally:
        eng.set_host_chunk(0)
    for u, n in enumerate(lengths):
        assert np.array_equal(y_plan[u, :n], ref[u, :n]), f"pipelined rb_process_host differs at utterance {u}"
        if algo in (0, 2):  # no transcendental on the path: device-drawn plans give the identical bits
            assert np.array_equal(y_seeded[u, :n], ref[u, :n]), f"seeded host entry differs at utterance {u}"
        else:               # taps / noise may differ by 1 ulp(fp32) from numpy's
            assert np.max(np.abs(y_seeded[u, :n].astype(np.float64) - ref[u, :n])) <= 1e-5


def test_seeded_host_entry_against_oracle(eng):
    """End to end against the float64 oracle: seeds in, waveforms out."""
    rs = np.random.RandomState(5)
    lengths = [64600, 16000, 64600, 2600]
    waves = [(0.3 * rs.standard_normal(n)).astype(np.float32) for n in lengths]
    seeds = [1234 + u for u in range(len(lengths))]
    ld = 64600
    xh = np.zeros((len(lengths), ld), np.float32)
    for u, w in enumerate(waves):
        xh[u, :w.shape[0]] = w
    y = eng.process_host_seeded(5, xh, np.array(lengths, np.int32), seeds, 16000, ARGS)
    for u, w in enumerate(waves):
        np.random.seed(seeds[u])
        ref = orc.process(w, 16000, ARGS, 5)
        assert np.max(np.abs(y[u, :w.shape[0]].astype(np.float64) - ref)) <= 1e-5


def test_device_plan_many_seeds_full_length(eng, P):
    """512 utterances of the benchmark's size: device-drawn integers and impulse gains against the native host planner
    (itself pinned to numpy by tests/test_host_logic.py), and against numpy itself on a subset."""
    from scl_deepfake_audio_detection_b200.native_planner import NativePlanner
    B = 512
    lengths = [64600] * B
    seeds = [(2654435761 * (u + 1)) % 2 ** 32 for u in range(B)]
    _, got, ld = device_plan(eng, lengths, seeds, 5)
    ref = NativePlanner(threads=0, pinned=False).draw(lengths, 16000, ARGS, 5, seeds=seeds, ld=ld, copy=True)
    assert np.array_equal(ref.lnl_tap_off, got.lnl_tap_off)
    assert np.array_equal(ref.isd_off, got.isd_off)
    assert np.array_equal(ref.isd_idx, got.isd_idx)
    assert np.array_equal(ref.isd_fr, got.isd_fr)
    assert ulp_err(got.lnl_taps, ref.lnl_taps) <= 1.0
    sub = P.draw_batch(lengths[:8], 16000, ARGS, 5, seeds=seeds[:8], ld=ld)
    n8 = int(sub.isd_off[-1])
    assert np.array_equal(sub.isd_off, got.isd_off[:9]) and np.array_equal(sub.isd_idx, got.isd_idx[:n8])
    assert np.array_equal(sub.isd_fr, got.isd_fr[:n8])
    assert np.array_equal(sub.lnl_tap_off, got.lnl_tap_off[:8 * 5 + 1])
    # every impulse position is a valid, unique sample index of its utterance
    for u in range(0, B, 37):
        idx = got.isd_idx[got.isd_off[u]:got.isd_off[u + 1]]
        assert idx.min(initial=0) >= 0 and idx.max(initial=0) < 64600 and np.unique(idx).size == idx.size


def test_streaming_submit_wait_equals_blocking_calls(eng):
    """rb_submit_host_seeded / rb_ctx_wait with two calls in flight deliver exactly what the blocking call delivers."""
    rs = np.random.RandomState(17)
    lengths = np.array([64600, 3000, 41234, 64600, 7, 20000], np.int32)
    ld = 64600
    xs = []
    for k in range(4):
        x = np.zeros((len(lengths), ld), np.float32)
        for u, n in enumerate(lengths):
            x[u, :n] = (0.4 * rs.standard_normal(n)).astype(np.float32)
        xs.append(x)
    seeds = [np.arange(100 * k, 100 * k + len(lengths), dtype=np.uint32) for k in range(4)]
    eng.set_host_chunk(2)
    try:
        ref = [eng.process_host_seeded(5, xs[k], lengths, seeds[k], 16000, ARGS) for k in range(4)]
        outs = [np.full_like(xs[0], np.nan) for _ in range(4)]
        tickets = []
        for k in range(4):
            tickets.append(eng.submit_host_seeded(5, xs[k], lengths, seeds[k], 16000, ARGS, out=outs[k]))
            if k >= 1:
                eng.wait_host(tickets[k - 1])
                for u, n in enumerate(lengths):
                    assert np.array_equal(outs[k - 1][u, :n], ref[k - 1][u, :n])
        eng.wait_host(0)
        for u, n in enumerate(lengths):
            assert np.array_equal(outs[3][u, :n], ref[3][u, :n])
        assert tickets == sorted(tickets) and len(set(tickets)) == 4
    finally:
        eng.set_host_chunk(0)
