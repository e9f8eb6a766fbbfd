for s in 100000 250000 500000 1000000 2000000; do
  for cfg in "8 22" "9 22" "10 22" "10 18" "11 18"; do
    set -- $cfg
    echo -n "S=$s maxlog2r=$1 large_min=$2: "
    RN_FFT_MAX_LOG2R=$1 RN_FFT_LARGE_TILE_MIN_LOG2L=$2 timeout 100 python tools/run_spectrum.py $s 20 | tail -1
  done
done
