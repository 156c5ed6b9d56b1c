"""First-contact probe for the tcgen05 conv kernel on a real B200 (debug tool, not product).

Runs a ladder of conv cases, each in its own subprocess with a timeout (a trapped kernel poisons the CUDA
context), in both UMMA descriptor conventions (hrnb_debug_set(0, swap)), and prints one line per case.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LADDER = [
    ("1x1 c64 mb1", (2, 16, 16, 64, 64, 1, 1, False, False, 1, False, False)),
    ("1x1 c16->16", (1, 8, 8, 16, 16, 1, 1, False, False, 1, False, False)),
    ("3x3 c32 mb1", (2, 64, 64, 32, 32, 3, 1, True, True, 1, False, False)),
    ("3x3 c32 mb2", (2, 64, 64, 32, 32, 3, 1, True, True, 2, False, False)),
    ("3x3 s2 gather", (2, 32, 32, 32, 64, 3, 2, False, False, 1, False, False)),
    ("1x1 gather", (2, 16, 16, 64, 64, 1, 1, False, False, 1, False, True)),
    ("3x3 c256 8x8", (4, 8, 8, 256, 256, 3, 1, True, True, 1, False, False)),
    ("final nchw", (2, 32, 32, 480, 21, 1, 1, False, False, None, True, False)),
]


def run_case(idx, swap):
    import torch
    from hrnet_b200 import _lib
    from test_gpu_kernels import _conv_case
    _lib.lib().hrnb_debug_set(0, swap)
    err, tol = _conv_case(*LADDER[idx][1])
    torch.cuda.synchronize()
    print(json.dumps({"case": LADDER[idx][0], "swap": swap, "err": err, "tol": tol, "ok": err <= tol}))


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "case":
        run_case(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    for swap in (0, 1):
        for i, (name, _) in enumerate(LADDER):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "case", str(i), str(swap)],
                                   capture_output=True, text=True, timeout=120)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                print(line[-1] if line else json.dumps({"case": name, "swap": swap, "rc": r.returncode,
                                                        "stderr": r.stderr[-600:]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"case": name, "swap": swap, "timeout": True}), flush=True)
