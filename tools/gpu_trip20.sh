#!/bin/bash
# grouped stride-2 backward (one wgrad + one dgrad launch per unit): tests + in-trip A/B (HRNB_S2_SPLIT=1 = per-phase launches)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_kernels.py -m gpu -q -x > $O/t20_pytest.txt 2>&1; echo "kernel tests rc=$?"; tail -3 $O/t20_pytest.txt
timeout 900 python -m pytest tests/test_gpu_train_network.py tests/test_gpu_network.py -m gpu -q -x > $O/t20_pytest_net.txt 2>&1; echo "network tests rc=$?"; tail -3 $O/t20_pytest_net.txt
for r in 1 2; do
for v in 0 1; do echo -n "train fuse_bwd_split=$v: "; HRNB_FUSE_BWD_SPLIT=$v timeout 600 python bench.py --no-cpu-baseline --no-infer 2>>$O/t20.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches']//d['steps'])"; done; done
tail -3 $O/t20.err
