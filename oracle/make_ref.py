"""Recipe for oracle/_ref: the UNMODIFIED reference implementation of the hot path, staged so it can travel to the
GPU box (TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package).

The reference is 100 % Python (there is nothing to compile), so "building" it means copying the files of the path
-- models/{model_internals,model_components,model_config1,model_config2}.py and Utils/{utils,EDM_sampler}.py --
byte for byte from /root/reference into oracle/_ref/, which is git-ignored (reference sources never enter the history)
but not gpurun-ignored.  bench.py's `--impl reference` arm and the `ref_cuda_eager` extra import it from there; when
the directory is absent they fall back to the committed oracle port (kind "port").

    python oracle/make_ref.py [/root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/model_internals.py", "models/model_components.py", "models/model_config1.py",
         "models/model_config2.py", "Utils/__init__.py", "Utils/utils.py", "Utils/EDM_sampler.py"]


def make(src="/root/reference") -> bool:
    if not os.path.isdir(src):
        return False
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": src, "sha256": manifest}, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES)


def import_ref():
    """(preconditioned_HDMOEM cfg1, preconditioned_HDMOEM cfg2, EDM_LOSS, MaskGenerator, EDM_Sampler) of the reference."""
    if not available():
        raise ImportError("oracle/_ref is absent: run `python oracle/make_ref.py` where /root/reference exists")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import importlib
    for name in ("models", "Utils"):          # the product package has no modules of these names; be safe anyway
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(DEST):
            del sys.modules[name]
    c1 = importlib.import_module("models.model_config1")
    c2 = importlib.import_module("models.model_config2")
    u = importlib.import_module("Utils.utils")
    s = importlib.import_module("Utils.EDM_sampler")
    return c1.preconditioned_HDMOEM, c2.preconditioned_HDMOEM, u.EDM_LOSS, u.MaskGenerator, s.EDM_Sampler


if __name__ == "__main__":
    ok = make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("oracle/_ref", "written" if ok else "NOT written (reference tree not found)")
