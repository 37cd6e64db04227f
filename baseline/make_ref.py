"""Stage the UNMODIFIED reference (Algolzw/image-restoration-sde) under baseline/_ref/ so that it travels to the GPU box.

The reference is plain Python with no package metadata: the prescribed
  python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
fails with "Neither 'setup.py' nor 'pyproject.toml' found" (recorded in DESIGN.md), so the files the sampling path needs
are copied verbatim instead - Python sources and YAMLs of `codes/utils`, `codes/data` and three task directories; no
checkpoints, images or notebooks.  baseline/_ref/ is git-ignored (never part of this repository's history) but NOT
gpurun-ignored.  Called by __graft_entry__.build() whenever /root/reference exists; a no-op on the GPU box.

Used only by: `bench.py --impl reference` (the reference's own IRSDE + ConditionalUNet, CPU or CUDA eager) and
tests/test_gpu_dropin.py (the reference's test.py run unchanged through `python -m irsde_b200.run`).
"""
import os
import shutil

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
TREES = ["codes/utils", "codes/data", "codes/config/deraining", "codes/config/denoising-sde", "codes/config/latent-dehazing"]
KEEP_EXT = (".py", ".yml", ".yaml", ".sh", ".txt")


def stage(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    stamp = os.path.join(DST, ".staged")
    if os.path.exists(stamp) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for tree in TREES:
        for root, dirs, files in os.walk(os.path.join(SRC, tree)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if not f.endswith(KEEP_EXT):
                    continue
                src = os.path.join(root, f)
                if os.path.getsize(src) > (1 << 20):
                    continue
                dst = os.path.join(DST, os.path.relpath(src, SRC))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                n += 1
    for extra in ("LICENSE", "requirements.txt"):
        if os.path.exists(os.path.join(SRC, extra)):
            shutil.copyfile(os.path.join(SRC, extra), os.path.join(DST, extra))
    commit = ""
    try:
        import json
        commit = json.load(open(os.path.join(SRC, ".SUBMODULES.json"))).get("commit", "")
    except Exception:
        pass
    with open(stamp, "w") as f:
        f.write("files=%d\nsource=%s\ncommit=%s\n" % (n, SRC, commit))
    return True


if __name__ == "__main__":
    print(stage(force=True), open(os.path.join(DST, ".staged")).read())
