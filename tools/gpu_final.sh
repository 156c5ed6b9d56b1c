#!/bin/bash
# round-end evidence trip: GPU suite, smoke, benches (product + reference arm), ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.txt 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/f_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.txt
timeout 400 python bench.py > gpurun_out/f_bench_train.json 2> gpurun_out/f_bench_train.err; echo "bench train rc=$?"; cut -c1-400 gpurun_out/f_bench_train.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/f_bench_reference.json
timeout 300 python bench.py --mode infer > gpurun_out/f_bench_infer_b64.json 2> gpurun_out/f_bench_infer.err; echo "bench infer rc=$?"; cut -c1-300 gpurun_out/f_bench_infer_b64.json
timeout 300 python bench.py --mode infer --batch 256 --no-cpu-baseline > gpurun_out/f_bench_infer_b256.json 2>> gpurun_out/f_bench_infer.err; echo "bench infer256 rc=$?"; cut -c1-300 gpurun_out/f_bench_infer_b256.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_train.csv python tools/train_once.py 64 2 > gpurun_out/f_ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/f_launches_train.csv > gpurun_out/f_launches_summary.txt 2>&1; head -12 gpurun_out/f_launches_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|wgrad_tc|bn_stats" --launch-skip 5 --launch-count 15 -o gpurun_out/f_prof_kernels python tools/kernel_once.py 64 > gpurun_out/f_ncu_kernels.log 2>&1; echo "ncu full rc=$?"
