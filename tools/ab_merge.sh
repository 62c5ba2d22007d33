#!/bin/bash
# tools/ab_merge.sh -- A/B of tuning builds on the one-to-one path: bench.py's merge block (1M read pairs, local + traceback) per build
for n in "$@"; do
  AMBIVERT_B200_LIB=ambivert_b200/align/tuning/lib_$n.so python bench.py --queries 3000 --no-glocal --no-cpu-baseline --no-pipeline --steps 3 > gpurun_out/bm_$n.json 2> gpurun_out/bm_$n.err
  python -c "
import json;d=json.loads(open('gpurun_out/bm_$n.json').read().strip().splitlines()[-1]);m=d['merge'];print('$n', round(m['kernel_ms'],2), round(m['gcups_kernel'],1), round(m['e2e_s']*1e3,1))" | tee -a gpurun_out/ab_merge.txt
done
