#!/bin/bash
# GPU trip 2: whole GPU suite, then A/B of the three new knobs on the default training bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t2_pytest.txt 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/t2_pytest.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline > gpurun_out/t2_$name.json 2> gpurun_out/t2_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t2_$name.json").read().strip().splitlines()[-1]); print("$name rc=$rc", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "serial", round(d["roofline"]["serial_step_ms"],2), d["roofline"]["share_of_serial_step"])
except Exception as e: print("$name rc=$rc parse fail", e)
PY
}
run all A=1
run all_again A=1
run no_wgstreams HRNB_WGRAD_STREAMS=0
run no_maskc HRNB_BN_MASK_C=0
run no_fuse HRNB_FUSE_STATS=0
run single HRNB_TRAIN_STREAMS=0
