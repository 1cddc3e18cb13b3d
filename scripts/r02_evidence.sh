#!/bin/bash
# round-2 evidence on one B200.  part A: bench lines; part B: ncu launch list + full capture of the dominant kernel.
# Only small files are left in gpurun_out/ (reports are converted to csv on the box).
part=${1:-A}
if [ "$part" = "A" ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
  timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
  timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu --no-other --modal --N 2048 --nl 10 > gpurun_out/r02_bench_c3_2048x10_modal.json 2>/dev/null
  timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu --no-other --N 1024 --nl 3 > gpurun_out/r02_bench_c2_1024x3.json 2>/dev/null
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-other --N 8192 --nl 4 > gpurun_out/r02_bench_8192x4_1gpu.json 2>/dev/null
  timeout 300 python bench.py --ensemble --steps 20 --warmup 5 > gpurun_out/r02_bench_c5_ensemble_1gpu.json 2>/dev/null
  timeout 300 python bench.py --ensemble --steps 10 --warmup 3 --noise libc --smoother lex > gpurun_out/r02_bench_c5_ensemble_1gpu_libc_lex.json 2>/dev/null
  for f in gpurun_out/r02_bench_*.json; do echo "$f: $(cut -c1-220 $f)"; done
else
  # launch list of the same command as the bench (cold-cache, serialised: shares only)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-other > gpurun_out/r02_ncu_bench.log 2>&1
  # full capture of the dominant kernel (finest-level k_relax_rb, 4 fused sweeps)
  ncu --set full --import-source on --clock-control none -k regex:k_relax_rb -s 6 -c 1 -o /tmp/r02_relax_rb python scripts/ncu_rb.py 4096 4 rb 1 > gpurun_out/r02_ncu_relax.log 2>&1
  ncu -i /tmp/r02_relax_rb.ncu-rep --page raw --csv > gpurun_out/r02_relax_rb.raw.csv 2>/dev/null
  ncu -i /tmp/r02_relax_rb.ncu-rep --page details --csv > gpurun_out/r02_relax_rb.details.csv 2>/dev/null
  # one instance of each of the other step kernels on the finest level (full set, raw page only)
  ncu --set full --clock-control none -k regex:"k_residual|k_correct|k_restrict|k_prolong4|k_lap2|k_rhs_t" -c 14 -o /tmp/r02_others python scripts/ncu_rb.py 4096 4 rb 1 > gpurun_out/r02_ncu_others.log 2>&1
  ncu -i /tmp/r02_others.ncu-rep --page raw --csv > gpurun_out/r02_others.raw.csv 2>/dev/null
  ls -la gpurun_out
fi
