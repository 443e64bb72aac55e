"""Per-utterance drop-in latency: process_Rawboost_feature(x, 16000, args, 5) as a loader would call it (one 64600-sample
utterance, host array in, host array out), with the draws by the native planner (default) and by numpy, next to the CPU oracle."""
import os, sys, time, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import rawboost_oracle as orc
from scl_deepfake_audio_detection_b200 import RawBoost as rb, workload

args = workload.default_args()
x = workload.synth_utterance(3, 64600, False)
def bench(fn, n):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
for planner in ("native", "numpy"):
    rb._PLANNER = planner
    np.random.seed(1)
    for algo in (5, 2, 3, 3, 2, 5):
        ms = bench(lambda: rb.process_Rawboost_feature(x, 16000, args, algo), 30)
        print(f"B200 drop-in, draws by {planner:6s} algo {algo}: {ms:7.3f} ms / utterance  ({1e3/ms:7.1f} utt/s per caller)")
np.random.seed(1)
for algo in (5, 2, 3):
    ms = bench(lambda: orc.process(x, 16000, orc.make_args(), algo), 5)
    print(f"CPU oracle (numpy/scipy)          algo {algo}: {ms:7.3f} ms / utterance  ({1e3/ms:7.1f} utt/s per core)")
