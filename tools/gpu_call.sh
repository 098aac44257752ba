#!/bin/bash
# scratch command script for one gpurun call
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "image_loss" 2>&1 | tail -5 > gpurun_out/loss_tests.txt
for v in "" loss_old loss_r36 loss_r54 loss_r90 loss_r108 loss_c64_r72 loss_c64_r36 loss_m3_r72; do
  if [ -z "$v" ]; then lib=""; else lib="gaussianhaircut_b200/lib/libgh_raster_$v.so"; fi
  echo "== ${v:-product}" >> gpurun_out/loss_variants.txt
  python tools/loss_case.py 200 $lib >> gpurun_out/loss_variants.txt 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/loss_launches.csv python tools/loss_case.py 3 > /dev/null 2>&1
cat gpurun_out/loss_tests.txt gpurun_out/loss_variants.txt
grep "gh_loss" gpurun_out/loss_launches.csv | tail -5
