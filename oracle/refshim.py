"""Import shim for the UNMODIFIED reference (xxlong0/NeuralUDF) -- test infrastructure only.

The reference lives read-only at /root/reference in the dev container; on the GPU box only the
byte-for-byte staged copy `oracle/_ref/` (oracle/make_ref.py; git-ignored, shipped with the snapshot)
exists.  This module is used by oracle/make_golden.py, by the `not gpu` tests that pin the oracle
restatement (oracle/oracle_torch.py) against the real reference code, and by bench.py's CPU-baseline /
`--impl reference` arm.  Nothing on the product path imports it.

The reference's `models/udf_renderer_blending.py:6-9` and `models/fields.py:6` import
`mcubes, icecream, skimage.measure, termcolor`, none of which is used on the render path; they are
replaced by empty stub modules here.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")          # byte-for-byte copies made by oracle/make_ref.py (git-ignored; travel to the GPU box)


def _pick_root():
    r = os.environ.get("NUDF_REFERENCE_ROOT", "/root/reference")
    if os.path.isfile(os.path.join(r, "models", "udf_renderer_blending.py")):
        return r
    return _STAGED


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "udf_renderer_blending.py"))


def is_staged_copy() -> bool:
    return os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_STAGED)


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def load():
    """Returns (fields_module, renderer_module) of the real reference."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _stub("mcubes")
    _stub("icecream", ic=lambda *a, **k: None)
    _stub("skimage")
    _stub("skimage.measure")
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    _stub("termcolor", colored=lambda s, *a, **k: s)
    # the reference's `models` is a namespace package rooted at REFERENCE_ROOT; make sure no other
    # `models` package (ours) shadows it inside this process.
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    fields = importlib.import_module("models.fields")
    renderer = importlib.import_module("models.udf_renderer_blending")
    assert fields.__file__.startswith(REFERENCE_ROOT), fields.__file__
    return fields, renderer
