#!/bin/bash
# N-GPU bench lines for the scaling tables: c3 (weak, full line with e2e) and c4 (strong, 10M frames in total)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "bench c3 exit $?"; tail -1 gpurun_out/bench_c3_n$N.json | cut -c1-300
timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload c4 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
echo "bench c4 exit $?"; tail -1 gpurun_out/bench_c4_n$N.json | cut -c1-300
