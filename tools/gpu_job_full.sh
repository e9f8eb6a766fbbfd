#!/bin/bash
# GPU job (1 GPU): all gpu tests, default bench, small workloads
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -8 gpurun_out/gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -1 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
for w in c1 c2; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -1 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err; done
