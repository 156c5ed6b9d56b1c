#!/bin/bash
# last check of a build: whole GPU suite, smoke, default bench (no ncu)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/final_pytest.txt 2>&1; echo "suite rc=$?"; tail -n 3 $O/final_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 2 $O/final_smoke.txt
timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('train', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],4), 'launches/step', d['gpu_launches']//d['steps'])
for i in d['infer']: print('infer', i['batch_per_gpu'], round(i['value']), round(i['tensor_frac_of_burst_peak'],4), 'e2e', round(i['e2e']['value']))
print('cpu_baseline', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
PY
