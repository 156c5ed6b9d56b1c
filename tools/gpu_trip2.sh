#!/bin/bash
# round 2, trip 2: full suite (no -x), default bench (PDL default), per-launch detail dumps
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > $O/t2_pytest.txt 2>&1; echo "suite rc=$?"; tail -25 $O/t2_pytest.txt
timeout 600 python bench.py --detail $O/t2_train_detail.json > $O/t2_bench.json 2> $O/t2_bench.err; echo "bench rc=$?"; cut -c1-300 $O/t2_bench.json; tail -3 $O/t2_bench.err
timeout 300 python bench.py --mode infer --batch 256 --no-cpu-baseline --detail $O/t2_infer256_detail.json > $O/t2_infer256.json 2>> $O/t2_bench.err; echo "infer256 rc=$?"; cut -c1-200 $O/t2_infer256.json
timeout 300 python bench.py --config 1 --no-cpu-baseline > $O/t2_config1.json 2>> $O/t2_bench.err; echo "config1 rc=$?"; cut -c1-300 $O/t2_config1.json
timeout 600 python bench.py --config 3 --sweep 1,2,4,8,16,32,64,128,256 --steps 10 --no-cpu-baseline > $O/t2_config3_sweep.json 2>> $O/t2_bench.err; echo "config3 rc=$?"; cut -c1-300 $O/t2_config3_sweep.json
timeout 600 python bench.py --config 5 --no-cpu-baseline --no-infer > $O/t2_config5.json 2>> $O/t2_bench.err; echo "config5 rc=$?"; cut -c1-300 $O/t2_config5.json
timeout 300 python bench.py --config 4 > $O/t2_config4.json 2>> $O/t2_bench.err; echo "config4 rc=$?"; cut -c1-300 $O/t2_config4.json
cp $O/parity_report.jsonl $O/t2_parity_report.jsonl 2>/dev/null
