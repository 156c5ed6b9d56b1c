#!/bin/bash
# pack_batch rewrite: tests + timing
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_network.py -m gpu -q > $O/t19_pytest.txt 2>&1; echo "tests rc=$?"; tail -3 $O/t19_pytest.txt
for r in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-infer --detail $O/t19_train_detail.json 2>>$O/t19.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('train', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"; done
python - <<'PY'
import json
L=json.load(open('gpurun_out/t19_train_detail.json'))['per_launch_ms']
for n,t in L:
    if n.split(':')[0] in ('pack','adam','repack','opt') or 'pack' in n or 'adam' in n: print(n, round(t*1e3,1))
print(len(L), sum(t for _,t in L))
PY
