#!/bin/bash
# final 1-GPU job of a round: all gpu tests, the single-GPU bench lines, launch lists and the affine kernel's ncu capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu_tests.log; tail -4 gpurun_out/gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -1 gpurun_out/bench_c3.json | cut -c1-300; tail -3 gpurun_out/bench_c3.err
for w in c1 c2; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -1 gpurun_out/bench_$w.json | cut -c1-300; tail -3 gpurun_out/bench_$w.err; done
python bench.py --workload c3dense --no-cpu-baseline --no-e2e > gpurun_out/bench_c3dense.json 2> gpurun_out/bench_c3dense.err; tail -1 gpurun_out/bench_c3dense.json | cut -c1-300
for w in c3 c2; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_launches_$w.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:affine_tma_kernel -s 3 -c 1 -f -o gpurun_out/r02_affine_c3 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_affine_c3.log 2>&1
ls -la gpurun_out/r02_affine_c3.ncu-rep gpurun_out/launches_c3.csv gpurun_out/launches_c2.csv
