"""GPU parity of the step after the path: batch_pad_for_multiview + view assembly on the device, against
outputs of the unmodified reference (tests/golden/multiview_golden.*, made by oracle/make_golden.py --multiview)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from conftest import GOLDEN_DIR, stream_digest  # noqa: E402
from oracle import rawboost_oracle as orc  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def mv():
    arrays = np.load(os.path.join(GOLDEN_DIR, "multiview_golden.npz"))
    with open(os.path.join(GOLDEN_DIR, "multiview_golden.json")) as f:
        return arrays, json.load(f)


@pytest.fixture(scope="module")
def eng():
    from scl_deepfake_audio_detection_b200.engine import Engine
    assert torch.cuda.is_available()
    return Engine(0)


def test_batch_pad_for_multiview_golden(eng, mv):
    """Every branch of the reference function, bit for bit (samples are copied, indices are integers), and the same draws."""
    from scl_deepfake_audio_detection_b200 import multiview
    arrays, meta = mv
    for key, m in meta["pad"].items():
        flat = arrays[m["input"]]
        views, o = [], 0
        for n in m["lens"]:
            views.append(flat[o:o + n].reshape(n, 1))
            o += n
        np.random.seed(m["seed"])
        out = multiview.batch_pad_for_multiview(views, 16000, m["length"], random_trim_nosil=m["trim"], repeat_pad=m["repeat_pad"])
        assert stream_digest() == m["stream"], key
        got = np.concatenate(out, axis=1)
        assert got.shape == arrays[key].shape, key
        assert np.array_equal(got, arrays[key].astype(np.float32)), key


@pytest.mark.parametrize("layout", [0, 1])
def test_assemble_layouts_and_batching(eng, layout):
    """Several items in one launch, both output layouts, against the oracle's index map."""
    from scl_deepfake_audio_detection_b200 import multiview
    rs = np.random.RandomState(3 + layout)
    G, V, length = 5, 4, 300
    lens = rs.randint(1, 700, size=(G, V))
    lens[1, 0] = 120   # first view shorter than the crop
    lens[3, 0] = 300   # exactly the crop
    waves = [rs.standard_normal(n).astype(np.float32) for n in lens.reshape(-1)]
    x, ln = eng.pack_waveforms(waves)
    for repeat_pad in (False, True):
        starts = [0 if lens[g, 0] < length else int(rs.randint(0, lens[g, 0] - length + 1)) for g in range(G)]
        out, out_len = multiview.assemble(eng, x, ln, V, starts, length, repeat_pad, layout)
        out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
        for g in range(G):
            first = int(lens[g, 0])
            wrap = first < length and repeat_pad
            olen = first if (first < length and not repeat_pad) else length
            assert out_len[g] == olen
            ref = orc.multiview_gather(waves[g * V:(g + 1) * V], first, starts[g], olen, wrap, repeat_pad)
            got = out[g, :olen, :] if layout == 0 else out[g, :, :olen].T
            assert np.array_equal(got, ref.astype(np.float32)), (g, repeat_pad)


def test_item_views_match_reference_item(eng, mv):
    """A whole Dataset item (RawBoost12 on 3 vocoded copies + the anchor, shared crop) in the loader's RNG order."""
    from scl_deepfake_audio_detection_b200 import multiview
    arrays, meta = mv
    args = orc.make_args()
    for key, m in meta["item"].items():
        waves = [orc.synth_utterance(m["first_wave"] + k, m["L"] + 37 * k, bool(k % 2)) for k in range(4)]
        np.random.seed(m["seed"])
        out, out_len = multiview.item_views(eng, [(waves[0], waves[1:])], args, 16000, m["trim"], repeat_pad=True,
                                            random_trim_nosil=True, layout=multiview.LAYOUT_ITEM)
        assert stream_digest() == m["stream"], "draw order / count differs from Dataset_for.__getitem__"
        got = out[0].cpu().numpy()
        assert list(got.shape) == m["shape"] and int(out_len[0]) == m["shape"][0]
        assert np.max(np.abs(got.astype(np.float64) - arrays[key])) <= 1e-5
        # the untouched views (anchor, vocoded copies) are pure copies
        for col in (0, 2, 3, 4):
            assert np.array_equal(got[:, col], arrays[key][:, col])


@pytest.mark.parametrize("L,K", [(64600, 8000), (16000, 1), (5000, 513), (300, 2049), (1, 700)])
def test_reverb_view_matches_reference_arithmetic(eng, L, K):
    """np.convolve(data, rir) + peak normalisation (audio_augmentor/reverb.py:39-42) on the FIR machinery."""
    from scl_deepfake_audio_detection_b200 import reverb
    rs = np.random.RandomState(L + K)
    x = (0.2 * rs.standard_normal(L)).astype(np.float32)
    rir = (rs.standard_normal(K) * np.exp(-np.arange(K) / max(1.0, K / 6.0))).astype(np.float32)
    y = reverb.reverb_convolve(x, rir)
    ref = orc.reverb_convolve(x, rir)
    assert y.dtype == np.float32 and y.shape == ref.shape == (L + K - 1,)
    assert np.max(np.abs(y.astype(np.float64) - ref)) <= 1e-5
    assert abs(float(np.max(np.abs(y))) - 1.0) <= 1e-6
