"""Install the B200 RawBoost path under the names the reference's loaders bind to.

The loaders do ``from datautils.RawBoost import ISD_additive_noise, LnL_convolutive_noise, SSI_additive_noise,
normWav`` at import (``/root/reference/datautils/asvspoof_2019_augall_3.py:10``) and define their own
``process_Rawboost_feature`` / ``RawBoost12`` that look those four names up in module globals at call time.
Two ways in, both leave the loader sources untouched:

* :func:`install` before the loader is imported: registers this package's ``RawBoost`` module as
  ``sys.modules['datautils.RawBoost']``.
* :func:`patch_loader` after import: rebinds the operator names (and the dispatcher, which then issues one fused
  device call per utterance instead of one per operator) inside an already imported loader module.

CUDA cannot be used in forked DataLoader workers once the parent initialised it (main.py:335 precedes main.py:379):
run the patched loaders with ``num_workers=0`` or ``multiprocessing_context='spawn'``, or use the batched path
(``plans.draw_batch`` in the workers, ``engine.Engine.process`` on the collated batch).
"""
from __future__ import annotations

import sys
import types

OPERATORS = ("randRange", "normWav", "genNotchCoeffs", "filterFIR", "LnL_convolutive_noise", "ISD_additive_noise",
             "SSI_additive_noise")
DISPATCH = ("process_Rawboost_feature", "RawBoost12")


def install(module_name: str = "datautils.RawBoost") -> types.ModuleType:
    """Make ``import datautils.RawBoost`` (or ``module_name``) resolve to the B200 implementation."""
    from . import RawBoost as impl
    sys.modules[module_name] = impl
    parent, _, leaf = module_name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, impl)
    return impl


def patch_loader(loader: types.ModuleType, dispatcher: bool = True) -> types.ModuleType:
    """Rebind the RawBoost names inside an imported loader module (e.g. ``datautils.asvspoof_2019_augall_3``)."""
    from . import RawBoost as impl
    for name in OPERATORS:
        if hasattr(loader, name):
            setattr(loader, name, getattr(impl, name))
    if dispatcher:
        for name in DISPATCH:
            if hasattr(loader, name):
                setattr(loader, name, getattr(impl, name))
    return loader
