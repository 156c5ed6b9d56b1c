#!/bin/bash
# fuse-layer up-path units: BatchNorm kernels batched per fuse output - tests + in-trip A/B
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_network.py -m gpu -q -x > $O/t24_pytest.txt 2>&1; echo "train network tests rc=$?"; tail -n 3 $O/t24_pytest.txt
for r in 1 2; do for v in 1 0; do echo -n "train fuse_bn_batch=$v: "; HRNB_FUSE_BN_BATCH=$v timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t24.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches']//d['steps'])"; done; done
tail -n 3 $O/t24.err
