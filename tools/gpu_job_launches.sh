#!/bin/bash
# launch lists of the library's own kernels for one bench run per workload (the synthetic-trajectory generation's
# torch kernels are filtered out by name), after the same command has run without ncu
mkdir -p gpurun_out
K='regex:affine_|dense_kernel|level_kernel|tile_kernel|pack_|combine_|fill_alpha0|smear_|apply_pbc|final_dist|finish_dist'
for w in c3 c2 c1; do
python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > /dev/null 2>&1 || echo "plain $w failed"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_launches_$w.log 2>&1
grep -c "," gpurun_out/launches_$w.csv
done
