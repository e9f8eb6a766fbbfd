#!/bin/bash
# 8-GPU baseline of round 2: host-to-device topology, per-phase breakdown, default bench line
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 tools/h2d_topology.py --mb 1024 > gpurun_out/h2d_topology_n$N.json 2> gpurun_out/h2d_topology_n$N.err
echo "topology exit $?"
timeout 300 $TR --master-port 29502 tools/dist_breakdown.py > gpurun_out/dist_breakdown_n$N.json 2> gpurun_out/dist_breakdown_n$N.err
echo "breakdown exit $?"; tail -1 gpurun_out/dist_breakdown_n$N.json
timeout 600 $TR --master-port 29503 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "bench exit $?"; tail -1 gpurun_out/bench_c3_n$N.json | cut -c1-1500
