#!/usr/bin/env python3
"""Recipe for ``oracle/_ref/``: the reference's own RawBoost code, compiled -- not copied -- from where it lies.

TEST INFRASTRUCTURE. The reference is pure Python (no C sources to build), so its "binary" is CPython bytecode:

* ``oracle/_ref/RawBoost.bytecode`` (a .pyc under a name snapshot tools do not filter out) <- ``py_compile`` of ``/root/reference/datautils/RawBoost.py`` (the seven operators);
* ``oracle/_ref/dispatch.marshal`` <- the code object of ``process_Rawboost_feature`` taken out of the compiled (never
  executed) loader module ``/root/reference/datautils/asvspoof_2019_augall_3.py:377-439`` -- the loader's module-level
  imports (librosa, soundfile, pydub, torchaudio ...) are therefore never needed.

``oracle/_ref/`` is git-ignored (no reference source or derived artefact enters the history) but not gpurun-ignored, so
the bytecode travels to the GPU box, where ``bench.py``'s CPU legs time the reference itself (``kind: "reference"``) and
``tests/`` pin the numpy restatement in ``rawboost_oracle.py`` against it. ``/root/reference`` only exists in the build
container; elsewhere :func:`build` is a no-op and the prebuilt files are used.
"""
import importlib.machinery
import importlib.util
import marshal
import os
import py_compile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference"
OPERATORS_SRC = os.path.join(REF, "datautils", "RawBoost.py")
LOADER_SRC = os.path.join(REF, "datautils", "asvspoof_2019_augall_3.py")


def build() -> bool:
    """Compile the reference into ``oracle/_ref/``. Returns False (and does nothing) where /root/reference is absent."""
    if not (os.path.isfile(OPERATORS_SRC) and os.path.isfile(LOADER_SRC)):
        return False
    os.makedirs(OUT, exist_ok=True)
    py_compile.compile(OPERATORS_SRC, cfile=os.path.join(OUT, "RawBoost.bytecode"), doraise=True,
                       invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(LOADER_SRC) as f:
        module_code = compile(f.read(), LOADER_SRC, "exec")
    found = [c for c in module_code.co_consts if isinstance(c, types.CodeType) and c.co_name == "process_Rawboost_feature"]
    if len(found) != 1:
        raise RuntimeError("process_Rawboost_feature not found in the reference loader")
    with open(os.path.join(OUT, "dispatch.marshal"), "wb") as f:
        f.write(importlib.util.MAGIC_NUMBER.hex().encode() + b"\n")  # bytecode is only valid for the interpreter that made it
        marshal.dump(found[0], f)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(OUT, "RawBoost.bytecode")) and os.path.isfile(os.path.join(OUT, "dispatch.marshal"))


_cache = None
last_error = None  # why load() returned None, for the bench's report


def load():
    """(operators module, process_Rawboost_feature) of the compiled reference, or None when ``oracle/_ref`` is missing or
    was built by another interpreter version."""
    global _cache, last_error
    if _cache is not None:
        return _cache
    if not available():
        last_error = f"{OUT} not present"
        return None
    try:
        loader = importlib.machinery.SourcelessFileLoader("_reference_RawBoost", os.path.join(OUT, "RawBoost.bytecode"))
        spec = importlib.util.spec_from_loader("_reference_RawBoost", loader)
        ops = importlib.util.module_from_spec(spec)
        loader.exec_module(ops)
        with open(os.path.join(OUT, "dispatch.marshal"), "rb") as f:
            if f.readline().rstrip(b"\n") != importlib.util.MAGIC_NUMBER.hex().encode():
                last_error = "oracle/_ref was compiled by another CPython bytecode version"
                return None
            code = marshal.load(f)
        # the dispatcher resolves the operator names in its module globals at call time: give it the operators
        dispatch = types.FunctionType(code, dict(vars(ops)), "process_Rawboost_feature")
    except Exception as e:  # noqa: BLE001 -- any failure means "use the port", and the reason is reported
        last_error = repr(e)
        return None
    _cache = (ops, dispatch)
    return _cache


if __name__ == "__main__":
    print("built" if build() else "reference not present; nothing built", OUT)
