"""Stage the UNMODIFIED reference for the GPU box -- test / baseline infrastructure only.

    python oracle/make_ref.py          # /root/reference -> oracle/_ref/   (dev container only)

The reference (xxlong0/NeuralUDF) is pure Python: there is nothing to compile, so the "build" of `oracle/_ref` is a
byte-for-byte copy of the files the hot path and its caller consist of.  `oracle/_ref/` is git-ignored (no reference
source enters the history) but travels to the GPU box with the gpurun snapshot like the built `libnudf.so`, where
`/root/reference` does not exist.  Consumers:

  * `bench.py --impl reference` and the `cpu_baseline` leg: time the reference's own `UDFRendererBlending.render_core`
    (models/udf_renderer_blending.py:327-584) on the host cores -> `cpu_baseline.kind = "reference"`;
  * `tests/test_runner_e2e.py`: runs the unmodified `exp_runner_blending.py` on top of the nudf modules.

A manifest with the sha256 of every staged file is written next to the copies; `verify()` re-checks it so that a stale or
edited copy is never silently used as "the reference".
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("NUDF_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")

# the render path ...
PATH_FILES = ["models/udf_renderer_blending.py", "models/fields.py", "models/embedder.py", "models/patch_projector.py",
              "models/projector_utils.py"]
# ... and its unmodified caller with what it imports (out-of-scope components that must run unchanged)
CALLER_FILES = ["exp_runner_blending.py", "extract_mesh.py", "dataset/dataset.py", "loss/__init__.py", "loss/loss.py",
                "loss/patch_metric.py", "confs/udf_dtu_blending.conf", "confs/udf_dtu_blending_ft.conf",
                "confs/udf_garment_blending.conf", "confs/udf_garment_blending_ft.conf"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    """Copies the files; returns the manifest dict.  No-op (returns None) when the reference checkout is absent."""
    if not os.path.isfile(os.path.join(SRC, PATH_FILES[0])):
        return None
    manifest = {}
    for rel in PATH_FILES + CALLER_FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "xxlong0/NeuralUDF (unmodified copies)", "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("staged %d reference files under %s" % (len(manifest), DST))
    return manifest


def available():
    return os.path.isfile(os.path.join(DST, "MANIFEST.json")) and os.path.isfile(os.path.join(DST, PATH_FILES[0]))


def verify():
    """True when every staged file still has the recorded hash."""
    if not available():
        return False
    man = json.load(open(os.path.join(DST, "MANIFEST.json")))["sha256"]
    return all(os.path.isfile(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in man.items())


if __name__ == "__main__":
    m = stage()
    if m is None:
        print("reference checkout not present at %s; nothing staged" % SRC)
        sys.exit(0 if available() else 1)
