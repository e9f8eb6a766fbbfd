#!/bin/bash
# 8-GPU job of round 2: parity test over all ranks, per-phase breakdown, c3 / c4 / c5 bench lines at their stated shapes
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k all_ranks > gpurun_out/multi_tests_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/multi_tests_n$N.log; tail -3 gpurun_out/multi_tests_n$N.log
timeout 300 $TR --master-port 29512 tools/dist_breakdown.py 1000000 > gpurun_out/dist_breakdown_n$N.json 2> gpurun_out/dist_breakdown_n$N.err
tail -1 gpurun_out/dist_breakdown_n$N.json
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "bench c3 exit $?"; tail -1 gpurun_out/bench_c3_n$N.json | cut -c1-400
timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload c4 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
echo "bench c4 exit $?"; tail -1 gpurun_out/bench_c4_n$N.json | cut -c1-400
timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err
echo "bench c5 exit $?"; tail -1 gpurun_out/bench_c5_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_c5_n$N.err
