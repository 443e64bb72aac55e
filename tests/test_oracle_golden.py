"""Pin the CPU oracle (oracle/rawboost_oracle.py) to the reference's own outputs.

The fixtures were produced by importing /root/reference unmodified (oracle/make_golden.py).
Plans / indices / stream positions must match bit-exactly; float64 waveforms to 1e-12.
"""
import numpy as np
import pytest

from conftest import stream_digest
from oracle import rawboost_oracle as orc

ARGS = orc.make_args()
TOL = 1e-12


def _case_keys(meta, section):
    return sorted(meta[section].keys())


def test_dispatcher_all_algos_match_reference(golden):
    arrays, meta = golden
    for key in _case_keys(meta, "cases"):
        c = meta["cases"][key]
        x = orc.synth_utterance(c["u"], c["L"], bool(c["loud"]))
        x_before = x.copy()
        np.random.seed(orc.seed_for(c["u"]))
        y = orc.process(x, 16000, ARGS, c["algo"])
        ref = arrays[key]
        assert np.asarray(y).dtype == ref.dtype, key
        assert np.asarray(y).shape == ref.shape, key
        assert np.array_equal(x, x_before), f"{key}: input mutated"
        assert (y is x) == c["same_object"], key
        np.testing.assert_allclose(y, ref, rtol=0, atol=TOL, err_msg=key)
        assert stream_digest() == c["stream"], f"{key}: RNG stream diverged"


def test_full_length_summaries_match_reference(golden):
    _, meta = golden
    for key in _case_keys(meta, "full"):
        s = meta["full"][key]
        algo = int(key.split("_")[0][4:])
        loud = int(key.split("_")[1][4:])
        u = int(key.split("_")[2][1:])
        x = orc.synth_utterance(u, 64600, bool(loud))
        np.random.seed(orc.seed_for(u))
        y = np.asarray(orc.process(x, 16000, ARGS, algo))
        assert str(y.dtype) == s["dtype"], key
        y = y.astype(np.float64)
        assert abs(y.sum() - s["sum"]) <= 1e-9, key
        assert abs((y * y).sum() - s["sumsq"]) <= 1e-9 * max(1.0, s["sumsq"]), key
        assert abs(y.min() - s["min"]) <= TOL and abs(y.max() - s["max"]) <= TOL, key
        assert int(np.argmax(np.abs(y))) == s["argmax_abs"], key
        np.testing.assert_allclose(y[np.array(s["probe_idx"])], s["probe_val"], rtol=0, atol=TOL)
        assert stream_digest() == s["stream"], key


def test_notch_taps_match_reference(golden):
    arrays, meta = golden
    for u in range(6):
        np.random.seed(orc.seed_for(u))
        b = orc.draw_notch_taps(5, 20, 8000, 100, 1000, 10, 100, 0, 0, 16000)
        ref = arrays[f"notch_u{u}"]
        assert b.shape[0] == meta["ops"][f"notch_u{u}"]["K"] and b.shape[0] % 2 == 1
        np.testing.assert_allclose(b, ref, rtol=0, atol=1e-15)
        assert stream_digest() == meta["ops"][f"notch_u{u}"]["stream"]
    np.random.seed(99)
    b = orc.draw_notch_taps(5, 20, 8000, 100, 1000, 10, 100, -5, -20, 16000)
    np.testing.assert_allclose(b, arrays["notch_gain"], rtol=0, atol=1e-15)
    assert stream_digest() == meta["ops"]["notch_gain"]["stream"]


@pytest.mark.parametrize("K", [1, 2, 3, 4, 11, 64, 491])
def test_filter_fir_matches_reference_and_closed_form(golden, K):
    arrays, _ = golden
    x, b, ref = arrays[f"fir_x_K{K}"], arrays[f"fir_b_K{K}"], arrays[f"fir_y_K{K}"]
    y = orc.filter_fir(x, b)
    assert y.dtype == ref.dtype and y.shape == ref.shape == x.shape
    np.testing.assert_allclose(y, ref, rtol=0, atol=TOL)
    np.testing.assert_allclose(orc.filter_fir_closed_form(x, b), ref, rtol=0, atol=1e-12)


def test_norm_wav_matches_reference(golden):
    arrays, _ = golden
    for tag in ("quiet", "loud"):
        x = arrays[f"norm_x_{tag}"]
        for always in (0, 1):
            y = orc.norm_wav(x, always)
            ref = arrays[f"norm_y{always}_{tag}"]
            assert y.dtype == ref.dtype
            assert np.array_equal(y, ref)
    x = arrays["norm_x_quiet"]
    assert orc.norm_wav(x, 0) is x  # untouched input is returned as the same object


def test_rand_range_matches_reference(golden):
    _, meta = golden
    r = meta["ops"]["randRange"]
    np.random.seed(3)
    f = orc.rand_range(20, 8000, 0)
    assert isinstance(f, np.ndarray) and f.shape == (1,) and float(f[0]) == r["float"]
    i = orc.rand_range(10, 100, 1)
    assert isinstance(i, int) and i == r["int"]
    assert float(orc.rand_range(-5, -20, 0)[0]) == r["reversed"]
    assert stream_digest() == r["stream"]


def test_operators_match_reference(golden):
    arrays, _ = golden
    for u in (0, 1):
        x = orc.synth_utterance(u, 16000, False)
        np.random.seed(orc.seed_for(u))
        np.testing.assert_allclose(orc.lnl(x, 5, 5, 20, 8000, 100, 1000, 10, 100, 0, 0, 5, 20, 16000),
                                   arrays[f"op_lnl_u{u}"], rtol=0, atol=TOL)
        np.random.seed(orc.seed_for(u))
        y = orc.isd(x, 10, 2)
        assert y.dtype == np.float32 and np.array_equal(y, arrays[f"op_isd_u{u}"])
        np.random.seed(orc.seed_for(u))
        np.testing.assert_allclose(orc.ssi(x, 10, 40, 5, 20, 8000, 100, 1000, 10, 100, 0, 0, 16000),
                                   arrays[f"op_ssi_u{u}"], rtol=0, atol=TOL)


def test_lnl_consumes_80_uniforms_and_isd_plan_is_integer_exact():
    np.random.seed(11)
    plan = orc.draw_lnl_plan(5, 5, 20, 8000, 100, 1000, 10, 100, 0, 0, 5, 20, 16000)
    after = stream_digest()
    np.random.seed(11)
    np.random.uniform(size=80)
    assert stream_digest() == after
    for t in plan.taps:
        assert 51 <= t.shape[0] <= 491 and t.shape[0] % 2 == 1
    np.random.seed(12)
    p = orc.draw_isd_plan(64600, 10)
    assert p.idx.dtype == np.int64 and len(np.unique(p.idx)) == len(p.idx) == int(64600 * (p.beta / 100))
    assert p.idx.min() >= 0 and p.idx.max() < 64600 and np.all(np.abs(p.f_r) < 1)


# ---------------------------------------------------------------------------------------------------------
# the step after the path: batch_pad_for_multiview and the item assembly
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mv_golden():
    import json
    import os
    from conftest import GOLDEN_DIR
    arrays = np.load(os.path.join(GOLDEN_DIR, "multiview_golden.npz"))
    with open(os.path.join(GOLDEN_DIR, "multiview_golden.json")) as f:
        return arrays, json.load(f)


def test_oracle_multiview_pad_matches_reference(mv_golden):
    from conftest import stream_digest
    arrays, meta = mv_golden
    assert len(meta["pad"]) == 32
    for key, m in meta["pad"].items():
        flat = arrays[m["input"]]
        views, o = [], 0
        for n in m["lens"]:
            views.append(flat[o:o + n].reshape(n, 1))
            o += n
        np.random.seed(m["seed"])
        out = orc.batch_pad_for_multiview(views, 16000, m["length"], random_trim_nosil=m["trim"], repeat_pad=m["repeat_pad"])
        out = np.concatenate(out, axis=1)
        assert out.shape == arrays[key].shape and out.shape[0] == m["out_len"], key
        assert np.array_equal(out.astype(np.float64), arrays[key].astype(np.float64)), key
        assert stream_digest() == m["stream"], f"{key}: the crop draw must consume the stream exactly like the reference"


def test_oracle_multiview_item_matches_reference(mv_golden):
    """RNG order of Dataset_for.__getitem__: RawBoost on the 3 vocoded copies, then on the anchor, then the shared crop."""
    from conftest import stream_digest
    arrays, meta = mv_golden
    args = orc.make_args()
    for key, m in meta["item"].items():
        item = int(key[4:])
        waves = [orc.synth_utterance(m["first_wave"] + k, m["L"] + 37 * k, bool(k % 2)) for k in range(4)]
        np.random.seed(m["seed"])
        aug_voc = [orc.process(w, 16000, args, 5) for w in waves[1:]]
        aug_anchor = orc.process(waves[0], 16000, args, 5)
        views = [waves[0], aug_anchor] + waves[1:] + aug_voc
        out = orc.batch_pad_for_multiview([np.expand_dims(v, 1) for v in views], 16000, m["trim"], random_trim_nosil=True,
                                          repeat_pad=True)
        out = np.concatenate(out, axis=1).astype(np.float32)
        assert list(out.shape) == m["shape"]
        assert np.max(np.abs(out.astype(np.float64) - arrays[key])) <= 1e-6, (key, item)
        assert stream_digest() == m["stream"]
