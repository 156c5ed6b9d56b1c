#!/bin/bash
# one GPU trip: new fused-statistics kernel tests, the whole GPU suite, A/B benches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -x -q -k "fused_bn" > gpurun_out/t1_fused.txt 2>&1; echo "fused tests rc=$?"; tail -3 gpurun_out/t1_fused.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t1_pytest.txt 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/t1_pytest.txt
for f in 1 0; do
  HRNB_FUSE_STATS=$f timeout 300 python bench.py --no-cpu-baseline > gpurun_out/t1_train_fuse$f.json 2> gpurun_out/t1_train_fuse$f.err; echo "train fuse=$f rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t1_train_fuse$f.json").read().strip().splitlines()[-1]); print("fuse=$f", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["roofline"]["share_of_serial_step"], d["roofline"]["serial_step_ms"])
except Exception as e: print("parse fail", e)
PY
done
for b in 64 128 256; do
  timeout 300 python bench.py --mode infer --batch $b --no-cpu-baseline > gpurun_out/t1_infer_b$b.json 2> gpurun_out/t1_infer_b$b.err; echo "infer b=$b rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t1_infer_b$b.json").read().strip().splitlines()[-1]); print("infer b=$b", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("parse fail", e)
PY
done
