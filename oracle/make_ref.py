"""Recipe that stages the UNMODIFIED reference renderer for the timed reference arms (test / measurement
infrastructure, never on the product path).

    python -m oracle.make_ref            # build container only (needs /root/reference)

Copies the three pure-Python files of the reference's ray-march, byte for byte, from where they lie under
``/root/reference/AvatarGen/AppearanceGen/models`` into ``oracle/_ref/models/``:

    embedder.py  fields.py  renderer.py

``oracle/_ref/`` is listed in ``.gitignore`` (the reference's sources never enter this repository's history) but
not in ``.gpurunignore``, so the staged copy travels to the GPU box exactly like the built ``.so`` does, and
``bench.py --impl reference`` / the ``ref_gpu`` leg of the native line can time the reference's own
``NeuSRenderer.render`` there (``/root/reference`` does not exist on that box).  A manifest with the sha256 of
every staged file is written next to them; ``load_reference_models()`` refuses files whose hash differs.

``renderer.py`` imports ``mcubes`` and ``icecream`` at module level (renderer.py:6-7); neither is used by
``render`` and neither is installable here, so ``load_reference_models`` puts empty stand-in modules into
``sys.modules`` before the import -- the files themselves are not edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import types

REF_MODELS = "/root/reference/AvatarGen/AppearanceGen/models"
HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("embedder.py", "fields.py", "renderer.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(quiet: bool = False) -> bool:
    """Copy the files (build container only).  Returns False when /root/reference is absent."""
    if not os.path.isdir(REF_MODELS):
        return False
    dst = os.path.join(REF_DIR, "models")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for name in FILES:
        shutil.copyfile(os.path.join(REF_MODELS, name), os.path.join(dst, name))
        manifest[name] = _sha(os.path.join(dst, name))
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_MODELS, "sha256": manifest}, f, indent=1)
    if not quiet:
        print(f"[make_ref] staged {', '.join(FILES)} -> {dst}")
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, "models", n)) for n in FILES) and \
        os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def load_reference_models():
    """Import the staged, unmodified ``models.fields`` / ``models.renderer`` (returns the two modules)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python -m oracle.make_ref` in the build container")
    man = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))["sha256"]
    for name in FILES:
        if _sha(os.path.join(REF_DIR, "models", name)) != man[name]:
            raise RuntimeError(f"oracle/_ref/models/{name} does not match its manifest hash")
    for name in ("mcubes", "icecream"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.ic = lambda *a, **k: None
            sys.modules[name] = m
    import importlib.util
    pkg = types.ModuleType("avc_refmodels")
    pkg.__path__ = [os.path.join(REF_DIR, "models")]
    sys.modules.setdefault("avc_refmodels", pkg)
    mods = {}
    for name in ("embedder", "fields", "renderer"):
        full = f"avc_refmodels.{name}"
        if full in sys.modules:
            mods[name] = sys.modules[full]
            continue
        spec = importlib.util.spec_from_file_location(full, os.path.join(REF_DIR, "models", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        # the files import each other as ``models.embedder``: alias the package name while they load
        prev = {k: sys.modules.get(k) for k in ("models", "models.embedder")}
        sys.modules["models"] = pkg
        if "embedder" in mods:
            sys.modules["models.embedder"] = mods["embedder"]
        try:
            spec.loader.exec_module(mod)
        finally:
            for k, v in prev.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        mods[name] = mod
    return mods["fields"], mods["renderer"]


if __name__ == "__main__":
    ok = stage()
    if not ok:
        raise SystemExit("reference not present: this recipe only runs in the build container")
