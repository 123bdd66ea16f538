"""Run the reference's UNMODIFIED `exp_runner_blending.py` on top of the nudf modules.

    python -m neuraludf_b200.launch /path/to/NeuralUDF/exp_runner_blending.py [runner args ...]

`models.fields`, `models.embedder`, `models.udf_renderer_blending` and `models.patch_projector` are bound to the nudf
mirrors before the runner is executed from the reference root; everything else the runner imports (`dataset`, `loss`, ...)
still resolves from the reference checkout (its `models/` directory is a namespace package).
"""
import importlib
import os
import runpy
import sys
import types


def install_shadow_modules(reference_root=None):
    """Registers the mirrors under the names the reference imports. Returns the `models` namespace module."""
    from neuraludf_b200.models import embedder, fields, patch_projector, udf_renderer_blending
    pkg = sys.modules.get("models")
    if pkg is None:
        pkg = types.ModuleType("models")
        pkg.__path__ = []
        sys.modules["models"] = pkg
    if reference_root is not None:
        p = os.path.join(reference_root, "models")
        if os.path.isdir(p) and p not in list(pkg.__path__):
            pkg.__path__ = list(pkg.__path__) + [p]
    for name, mod in (("fields", fields), ("embedder", embedder), ("udf_renderer_blending", udf_renderer_blending),
                      ("patch_projector", patch_projector)):
        sys.modules["models." + name] = mod
        setattr(pkg, name, mod)
    return pkg


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    runner = os.path.abspath(argv[0])
    root = os.path.dirname(runner)
    install_shadow_modules(root)
    if root not in sys.path:
        sys.path.insert(1, root)
    os.chdir(root)
    sys.argv = [runner] + argv[1:]
    runpy.run_path(runner, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
