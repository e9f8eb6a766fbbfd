#!/bin/bash
# ncu --set full of dense_kernel_tp on c2 (STO cubic, 8-byte copy path) and on LLZO forced dense (16-byte path)
mkdir -p gpurun_out
python tools/run_dense.py STO cubic 100000 0 > /dev/null 2>&1 || echo "plain run failed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_kernel_tp -s 2 -c 1 -f -o gpurun_out/r02_dense_c2_v2 \
    python tools/run_dense.py STO cubic 100000 0 > gpurun_out/ncu_dense_c2_v2.log 2>&1; tail -1 gpurun_out/ncu_dense_c2_v2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_kernel_tp -s 2 -c 1 -f -o gpurun_out/r02_dense_llzo \
    python tools/run_dense.py LLZO art 200000 1 > gpurun_out/ncu_dense_llzo.log 2>&1; tail -1 gpurun_out/ncu_dense_llzo.log
ls -la gpurun_out/r02_dense_c2_v2.ncu-rep gpurun_out/r02_dense_llzo.ncu-rep
