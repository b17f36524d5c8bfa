set -x
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 500 python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/bench_ref_1.json 2> gpurun_out/bench_ref_err.log; tail -c 600 gpurun_out/bench_ref_1.json
timeout 400 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_err.log; tail -c 900 gpurun_out/bench_ours_1.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 200 python tools/trace_mega.py > gpurun_out/trace_mega.txt 2>&1; tail -24 gpurun_out/trace_mega.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:megaDecodeKernel -s 2 -c 1 -f -o gpurun_out/prof_mega python tools/profile_targets.py > gpurun_out/ncu_mega.log 2>&1; tail -3 gpurun_out/ncu_mega.log
