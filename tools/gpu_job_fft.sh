#!/bin/bash
# GPU job: spectrum parity tests + timings + launch list of the spectrum path
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spectrum.py -x -q -m gpu > gpurun_out/fft_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/fft_tests.log
tail -15 gpurun_out/fft_tests.log
for s in 10000 100000 1000000 2000000 4000000 8000000; do python tools/run_spectrum.py $s 20; done > gpurun_out/fft_times.log 2>&1
cat gpurun_out/fft_times.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/fft_launches.csv python tools/run_spectrum.py 1000000 2 > gpurun_out/fft_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/fft_launches.csv pack_alpha 2>&1 | tail -20
