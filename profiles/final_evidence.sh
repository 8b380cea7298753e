#!/bin/bash
# one gpurun call that regenerates the round's evidence under gpurun_out/ (copied into profiles/ afterwards)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 1500 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-others > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ssd_fclk_mom -s 3 -c 1 -f -o gpurun_out/prof_mom python bench.py --one-arm --no-others --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ncc_update_f32 -s 3 -c 1 -f -o gpurun_out/prof_ncc_f32 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -n 12
