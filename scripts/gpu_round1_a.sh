#!/bin/bash
# first GPU pass of the session: parity tests, smoke, bench, launch list, full ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err
timeout 600 python bench.py --workload ensemble > gpurun_out/bench_ensemble.json 2> gpurun_out/bench_ensemble.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_sweep.csv python scripts/prof_driver.py sweep 1184 2 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv \
   --log-file gpurun_out/launches_design.csv python scripts/prof_driver.py design 16 2 >> gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'potrf_lnlike|write_kernel|moments_k|marginal' -c 10 \
   -o gpurun_out/prof_sweep -f python scripts/prof_driver.py sweep 592 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'design_kernel' -c 1 \
   -o gpurun_out/prof_design -f python scripts/prof_driver.py design 16 1 >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench_sweep.json
