"""B200-native RawBoost waveform augmentation (the hot path of josebeo2016/SCL-Deepfake-audio-detection).

Python host side over a hand-written sm_100a CUDA library with a C ABI (``include/rawboost_b200.h``):

* ``RawBoost``   -- drop-in for the reference's ``datautils/RawBoost.py`` operator surface plus
                    ``process_Rawboost_feature`` / ``RawBoost12`` (same names, arguments, RNG stream).
* ``plans``      -- host-side random parameters, drawn with the reference's own numpy calls in its order.
* ``engine``     -- batched device execution through the C ABI (torch only for memory and streams).
* ``dropin``     -- installs the replacement under the reference's module names.

There is no CPU fallback: importing works anywhere, computing needs the built library and a B200.
"""
__version__ = "0.1.0"
