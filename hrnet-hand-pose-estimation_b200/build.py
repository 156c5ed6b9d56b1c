"""Build libhrnb.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

Usage: python hrnet-hand-pose-estimation_b200/build.py [--force]
nvcc cross-compiles for sm_100a without a GPU, so this runs in the authoring container and the
resulting .so travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhrnb.so")
STAMP = os.path.join(HERE, ".libhrnb.stamp")
SOURCES = ["api.cu", "conv_tc.cu", "wgrad_tc.cu", "elementwise.cu", "train_ops.cu", "decode_loss.cu", "triangulate.cu", "glue.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


if os.environ.get("HRNB_HANG_RECORDS", "0") == "1":      # debug build: stuck warps leave records before the bounded-wait trap
    NVCC_FLAGS.append("-DHRNB_HANG_RECORDS")


def _digest():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/hrnb.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile when sources changed; returns the library path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


def build_variant(name, defines):
    """experiment builds (tools/): the same sources with extra -D flags into hrnet-hand-pose-estimation_b200/<name>, selected
    at run time with HRNB_LIB=<name>"""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(HERE, name)
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
