"""Recipe that vendors the UNMODIFIED reference into `oracle/_ref/` (TEST / BENCHMARK INFRASTRUCTURE ONLY).

The reference (IamJerryXu/Multimodal-Diagnosis-HAM-Spine) is 100 % Python: there is nothing to compile and no
setup.py / pyproject.toml to `pip install`.  So the "build" of the reference is a verbatim copy of the source files of the
hot path from where they lie under /root/reference into the git-ignored `oracle/_ref/` (listed in .gitignore, NOT in
.gpurunignore, so it travels to the GPU box like our own built .so files).  Nothing under `oracle/_ref/` is ever committed,
imported by the product package, or edited: `bench.py --impl reference` / `--impl torch_gpu` and the oracle-pinning tests
import it to time / compare against the real thing (`cpu_baseline.kind = "reference"`).

    python oracle/make_ref.py            # copy (no-op when /root/reference is absent, e.g. on the GPU box)
    python oracle/make_ref.py --check    # exit 1 when oracle/_ref is missing or stale

`__graft_entry__.build()` runs this whenever /root/reference is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("MDHS_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

# the files of SURVEY.md section 8a (hot path) plus what they import
FILES = [
    "model.py", "encoder.py",
    "modules/__init__.py", "modules/fusion_blocks.py", "modules/heads.py", "modules/gating.py", "modules/tabular.py",
    "modules/sequence_blocks.py",
    "mibf_net/__init__.py", "mibf_net/model_resnet.py", "mibf_net/attention.py", "mibf_net/bert.py",
    "ConNexT/models/ourmodel.py", "ConNexT/models/BERT.py", "ConNexT/models/block/moe.py", "ConNexT/models/block/kan1.py",
    "scripts/__init__.py", "scripts/train.py", "scripts/predict.py",
]


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def available():
    """True when a vendored reference is importable from oracle/_ref."""
    return os.path.exists(os.path.join(REF_DST, "model.py")) and os.path.exists(os.path.join(REF_DST, "MANIFEST.json"))


def make(verbose=False):
    if not os.path.isdir(REF_SRC):
        if verbose:
            print(f"make_ref: {REF_SRC} not present (GPU box): keeping {'existing' if available() else 'no'} oracle/_ref")
        return available()
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(src)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"make_ref: vendored {len(manifest)} reference files into {REF_DST}")
    return True


def check():
    if not available():
        return False
    man = json.load(open(os.path.join(REF_DST, "MANIFEST.json")))["files"]
    return all(os.path.exists(os.path.join(REF_DST, rel)) and _sha(os.path.join(REF_DST, rel)) == h for rel, h in man.items())


if __name__ == "__main__":
    if "--check" in sys.argv:
        sys.exit(0 if check() else 1)
    make(verbose=True)
