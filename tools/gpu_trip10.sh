#!/bin/bash
# round 2 evidence trip: whole GPU suite, smoke, default bench (+ reference arm), every BASELINE config, decode bench,
# ncu launch lists (training step, inference pass) and `--set full` captures of the dominant kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > $O/t10_pytest.txt 2>&1; echo "suite rc=$?"; tail -4 $O/t10_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t10_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/t10_smoke.txt
timeout 900 python bench.py > $O/t10_bench_default.json 2> $O/t10_bench_default.err; echo "bench default rc=$?"; cut -c1-600 $O/t10_bench_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/t10_bench_reference.json 2> $O/t10_bench_reference.err; echo "bench ref rc=$?"; cut -c1-400 $O/t10_bench_reference.json
for c in 1 3 4 5; do timeout 900 python bench.py --config $c --no-cpu-baseline > $O/t10_bench_config$c.json 2> $O/t10_bench_config$c.err; echo "config $c rc=$?"; cut -c1-300 $O/t10_bench_config$c.json; done
timeout 200 python tools/decode_bench.py 256 20 > $O/t10_decode_b256.jsonl 2>&1; tail -1 $O/t10_decode_b256.jsonl | cut -c1-300
# launch lists (serialised, cold cache: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/t10_launches_train.csv python tools/train_once.py 64 2 > $O/t10_ncu_launches_train.log 2>&1; echo "ncu train list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/t10_launches_infer.csv python tools/run_once.py 256 2 > $O/t10_ncu_launches_infer.log 2>&1; echo "ncu infer list rc=$?"
# full captures of the dominant kernels in isolation
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|wgrad_tc|bn_stats" --launch-skip 5 --launch-count 15 -f -o $O/t10_prof_train_kernels python tools/kernel_once.py 64 > $O/t10_ncu_train_kernels.log 2>&1; echo "ncu train full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|softmax_softargmax|decode_argmax|softargmax|final_preds|heatmap_loss" --launch-count 30 -f -o $O/t10_prof_infer_kernels python tools/kernel_once_infer.py 256 > $O/t10_ncu_infer_kernels.log 2>&1; echo "ncu infer full rc=$?"
for n in train infer; do ncu -i $O/t10_prof_${n}_kernels.ncu-rep --page raw --csv > $O/t10_prof_${n}_kernels_raw.csv 2>/dev/null; ncu -i $O/t10_prof_${n}_kernels.ncu-rep --page details --csv > $O/t10_prof_${n}_kernels_details.csv 2>/dev/null; done
ls -la $O/*.ncu-rep; rm -f $O/*.ncu-rep; du -sh $O
