#!/bin/bash
# N-GPU job: multi-GPU parity tests, per-phase breakdown, default bench line (no e2e)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/multi_tests_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/multi_tests_n$N.log; tail -4 gpurun_out/multi_tests_n$N.log
timeout 300 $TR --master-port 29512 tools/dist_breakdown.py 1000000 > gpurun_out/dist_breakdown_n$N.json 2> gpurun_out/dist_breakdown_n$N.err
tail -1 gpurun_out/dist_breakdown_n$N.json; tail -3 gpurun_out/dist_breakdown_n$N.err
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "bench exit $?"; tail -1 gpurun_out/bench_c3_n$N.json | cut -c1-2800; tail -4 gpurun_out/bench_c3_n$N.err
for ov in 0 1; do
RN_DIST_OVERLAP=$ov timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --no-e2e --no-parity > gpurun_out/bench_c3_n${N}_overlap$ov.json 2> gpurun_out/bench_c3_n${N}_overlap$ov.err
echo "overlap=$ov exit $?"; tail -1 gpurun_out/bench_c3_n${N}_overlap$ov.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['stages'])"
done
