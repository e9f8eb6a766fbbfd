#!/bin/bash
# GPU job (N GPUs): multi-GPU parity + bench lines.  usage: gpu_job_multi.sh N [extra workloads]
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/multi_tests_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/multi_tests_n$N.log
tail -25 gpurun_out/multi_tests_n$N.log
run() {  # workload, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $1 $2 > gpurun_out/bench_$1_n$N.json 2> gpurun_out/bench_$1_n$N.err
  echo "bench $1 exit $?"; tail -1 gpurun_out/bench_$1_n$N.json | cut -c1-3000; tail -5 gpurun_out/bench_$1_n$N.err
}
run c3 "--steps 20 --warmup 3"
shift
for w in "$@"; do run $w "--steps 10 --warmup 3"; done
