#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_install.py -x -q -m gpu > gpurun_out/n${N}_single_tests.log 2>&1; tail -5 gpurun_out/n${N}_single_tests.log
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/multi_tests_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/multi_tests_n$N.log; tail -6 gpurun_out/multi_tests_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_breakdown.py 1000000 > gpurun_out/dist_breakdown_n$N.json 2> gpurun_out/dist_breakdown_n$N.err
tail -1 gpurun_out/dist_breakdown_n$N.json; tail -3 gpurun_out/dist_breakdown_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "bench exit $?"; tail -1 gpurun_out/bench_c3_n$N.json | cut -c1-2500; tail -4 gpurun_out/bench_c3_n$N.err
