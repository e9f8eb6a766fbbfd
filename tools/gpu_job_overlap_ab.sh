#!/bin/bash
# N-GPU A/B of the overlapped schedule (RN_DIST_OVERLAP=0/1), after the single-GPU phase tests
N=${1:-4}
shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_polarizability.py -x -q -m gpu -k "two_phases" 2>&1 | tail -2
for cfg in "0 2" "0 16" "1 2" "1 16" "0 8"; do
set -- $cfg; ov=$1; ring=$2
RN_AFFINE_OUT_RING=$ring RN_DIST_OVERLAP=$ov timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --no-e2e --no-parity > gpurun_out/bench_c3_n${N}_overlap$ov.json 2> gpurun_out/bench_c3_n${N}_overlap$ov.err
echo "overlap=$ov ring=$ring exit $?"; tail -1 gpurun_out/bench_c3_n${N}_overlap$ov.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['stages'])"
done
RN_DIST_OVERLAP=1 timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
