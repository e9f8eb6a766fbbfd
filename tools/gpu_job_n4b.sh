#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/multi_tests_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/multi_tests_n$N.log; tail -4 gpurun_out/multi_tests_n$N.log
for pipe in 1 0; do
RN_DIST_PIPELINE=$pipe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 3 --no-e2e --no-parity > gpurun_out/bench_c3_n${N}_pipe$pipe.json 2> gpurun_out/bench_c3_n${N}_pipe$pipe.err
echo "bench pipe=$pipe exit $?"; tail -1 gpurun_out/bench_c3_n${N}_pipe$pipe.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['stages'])"; tail -2 gpurun_out/bench_c3_n${N}_pipe$pipe.err
done
