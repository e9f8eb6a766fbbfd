#!/bin/bash
# 1-GPU job: all gpu tests, the three single-GPU bench lines, ncu of the c2 dense kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectrum.py -x -q -m gpu > gpurun_out/spectrum_tests.log 2>&1
echo "spectrum pytest exit $?"; tail -4 gpurun_out/spectrum_tests.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu_tests.log; tail -6 gpurun_out/gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -1 gpurun_out/bench_c3.json | cut -c1-600; tail -3 gpurun_out/bench_c3.err
for w in c1 c2; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -1 gpurun_out/bench_$w.json | cut -c1-900; tail -3 gpurun_out/bench_$w.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_kernel_tp -s 4 -c 1 -f -o gpurun_out/r02_dense_c2 \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_dense_c2.log 2>&1
tail -2 gpurun_out/ncu_dense_c2.log | cut -c1-300
ls -la gpurun_out/r02_dense_c2.ncu-rep
