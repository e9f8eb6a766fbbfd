#!/bin/bash
# 1-GPU tuning job: FFT kernel variants, smearing timing, ncu of the spectrum kernels
mkdir -p gpurun_out
for lean in 0 1 2 3; do
  for s in 1000000 8000000; do echo -n "RN_FFT_LEAN=$lean "; RN_FFT_LEAN=$lean python tools/run_spectrum.py $s 30; done
done > gpurun_out/fft_lean.log 2>&1
cat gpurun_out/fft_lean.log
python tools/run_convolve.py > gpurun_out/convolve_times.log 2>&1; cat gpurun_out/convolve_times.log
ncu --set full --clock-control none --import-source on -k regex:"level_kernel|tile_kernel|pack_alpha|combine_md" -c 5 -s 10 \
    -o gpurun_out/r02_fft_full -f python tools/run_spectrum.py 1000000 2 > gpurun_out/ncu_fft_full.log 2>&1
tail -3 gpurun_out/ncu_fft_full.log
ls -la gpurun_out/*.ncu-rep
