"""Recipe for oracle/_ref/: an UNMODIFIED copy of the reference's own Python implementation of the hot path, placed
next to the oracle so that it travels to the GPU box (which has no /root/reference) and `bench.py --impl reference`
can time the real reference on the box's host cores instead of the oracle port (cpu_baseline.kind = "reference").
TEST / MEASUREMENT INFRASTRUCTURE - nothing in the product path imports it.

    python oracle/build_ref.py          (run by __graft_entry__.build() when /root/reference is present)

oracle/_ref/ is git-ignored (the repository never holds reference sources) but NOT gpurun-ignored.  Copied verbatim:
lib/{models,core,utils,config} (the pose_hrnet* modules the path uses import their siblings through models/__init__.py)
and the experiment YAMLs the benchmarks load.  A MANIFEST with the sha256 of every file is written beside them; the
loader (oracle/ref_shim.py with HRNB_REFERENCE_ROOT=oracle/_ref) applies the same import shims as for /root/reference.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("HRNB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
DIRS = ["lib/models", "lib/core", "lib/utils", "lib/config"]
FILES = ["lib/__init__.py",
         "experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml",
         "experiments/RHD/RHD_HRNet_w32_max_hmloss_v1.yaml",
         "experiments/RHD/RHD_HRNet_w48_trainable_softmax_hm-pose2dloss_v1.yaml"]


def build():
    if not os.path.isdir(os.path.join(SRC, "lib", "models")):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = []
    for d in DIRS:
        for root, _, files in os.walk(os.path.join(SRC, d)):
            for f in files:
                if f.endswith((".py", ".yaml")):
                    FILES.append(os.path.relpath(os.path.join(root, f), SRC))
    for rel in sorted(set(FILES)):
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as fh:
            manifest.append("%s  %s" % (hashlib.sha256(fh.read()).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref:", "built from " + SRC if ok else "reference not present, nothing built")
    sys.exit(0)
