#!/bin/bash
# Per-kernel ncu --set full captures of the finest-level launch of every hot-path kernel (one timestep of 4096^2 x 4).
# The reports embed the whole cubin (~50 MB each), so only the exported raw/details pages are kept.
# usage (on the GPU box): bash scripts/ncu_kernels.sh gpurun_out/ncu_r01
out=${1:-gpurun_out/ncu}
mkdir -p $out
cap() { # name regex skip count
  ncu --set full --clock-control none -k regex:"$2" --launch-skip $3 --launch-count $4 -f -o /tmp/$1 python scripts/one_step.py > $out/$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > $out/$1.details.txt 2>/dev/null
  rm -f /tmp/$1.ncu-rep
  tail -1 $out/$1.log
}
cap lap      '^k_lap2$'     0 1
cap rhs      '^k_rhs_t'     0 1
cap residual '^k_residual$' 0 1
cap correct  '^k_correct$'  0 1
cap restrict '^k_restrict$' 0 1
cap prolong  '^k_prolong4$' 10 1
cap relax    '^k_relax_ws'  11 1
du -sh $out
