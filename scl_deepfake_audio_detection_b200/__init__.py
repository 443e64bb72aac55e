"""Import name of the package whose sources live in ``scl-deepfake-audio-detection_b200/``.

The directory name the project mandates contains hyphens, which Python cannot import; this shim points the
package search path at it, so ``import scl_deepfake_audio_detection_b200.RawBoost`` loads
``scl-deepfake-audio-detection_b200/RawBoost.py``. No code lives here.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "scl-deepfake-audio-detection_b200")
if not _os.path.isdir(_SRC):  # pragma: no cover
    raise ImportError(f"package sources not found at {_SRC}")
__path__ = [_SRC]
with open(_os.path.join(_SRC, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_SRC, "__init__.py"), "exec"))
del _os, _f
