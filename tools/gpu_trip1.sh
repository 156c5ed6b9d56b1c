#!/bin/bash
# round 2, trip 1: verify everything staged blind (tests, new bench line), PDL soak, 2-CTA/SM experiment
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/t1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/t1_pytest.txt 2>&1; echo "suite rc=$?"; tail -15 $O/t1_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t1_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/t1_smoke.txt
timeout 600 python bench.py > $O/t1_bench.json 2> $O/t1_bench.err; echo "bench rc=$?"; cut -c1-600 $O/t1_bench.json; tail -3 $O/t1_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/t1_bench_ref.json 2> $O/t1_bench_ref.err; echo "ref rc=$?"; cut -c1-400 $O/t1_bench_ref.json
brief() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [ (i['batch_per_gpu'], round(i['value'])) for i in d.get('infer',[])])
except Exception as e: print('FAILED', e)
"; }
for i in 1 2 3 4 5 6 7 8; do echo -n "pdl train #$i: "; HRNB_TRAIN_PDL=1 HRNB_PDL=1 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --infer-batches 64 2>>$O/t1_pdl.err | brief; done
for i in 1 2 3; do echo -n "pdl infer256 #$i: "; HRNB_PDL=1 timeout 200 python bench.py --mode infer --batch 256 --no-cpu-baseline 2>>$O/t1_pdl.err | brief; done
echo -n "infer256 default: "; timeout 200 python bench.py --mode infer --batch 256 --no-cpu-baseline 2>>$O/t1_misc.err | brief
echo -n "infer64 epi8 share MB=2: "; HRNB_LIB=libhrnb_epi8.so HRNB_TMEM_SHARE=1 HRNB_MB=2 timeout 120 python bench.py --mode infer --batch 64 --no-cpu-baseline 2>>$O/t1_misc.err | brief
echo -n "infer64 epi8 share: "; HRNB_LIB=libhrnb_epi8.so HRNB_TMEM_SHARE=1 timeout 120 python bench.py --mode infer --batch 64 --no-cpu-baseline 2>>$O/t1_misc.err | brief
echo -n "infer64 epi8: "; HRNB_LIB=libhrnb_epi8.so timeout 120 python bench.py --mode infer --batch 64 --no-cpu-baseline 2>>$O/t1_misc.err | brief
echo -n "infer64 default: "; timeout 120 python bench.py --mode infer --batch 64 --no-cpu-baseline 2>>$O/t1_misc.err | brief
HRNB_TRAIN_PDL=1 HRNB_PDL=1 timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py -k "not conditioned and not trajectory" > $O/t1_pytest_pdl.txt 2>&1; echo "suite(PDL) rc=$?"; tail -3 $O/t1_pytest_pdl.txt
cp $O/parity_report.jsonl $O/t1_parity_report.jsonl 2>/dev/null
