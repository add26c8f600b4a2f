#!/bin/bash
# Round evidence: run on the GPU box through gpurun; everything lands in gpurun_out/ (copy the summaries to profiles/).
#   gpurun --timeout 1700 -- 'bash tools/capture_profiles.sh r02'
set -x
R=${1:-r02}
python tools/fp64_peak.py > gpurun_out/${R}_fp64_peak.json 2> /dev/null
python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference > gpurun_out/${R}_bench_reference_n1.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/${R}_plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_cfg3_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/${R}_ncu_launches.log 2>&1
for spec in "am500 fdm_lean_kernel 20000" "am500div fdm_lean_kernel 5000" "am400 fdm_lean_kernel 20000" "bar400 fdm_lean_kernel 20000" "cfg4 fdm_cta_kernel<\(int\)16,.\(int\)16 148" "dbfp fdm_cta_kernel<\(int\)14,.\(int\)5 1184" "p2 fdm_rowvar_kernel<\(int\)19,.\(bool\)1 3552"; do
  set -- $spec
  python tools/ncu_shapes.py $1 $3 > gpurun_out/${R}_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 1 -c 1 \
      -o gpurun_out/${R}_prof_$1 -f python tools/ncu_shapes.py $1 $3 > gpurun_out/${R}_ncu_$1.log 2>&1
  # the reports are ~23 MB each and gpurun brings back 64 MiB at most: summarise here, keep the text
  python tools/ncu_summary.py gpurun_out/${R}_prof_$1.ncu-rep gpurun_out/${R}_$1_ncu_full_summary.txt $3 > /dev/null
  python tools/ncu_opmix.py gpurun_out/${R}_prof_$1.ncu-rep > gpurun_out/${R}_$1_ncu_opmix.txt 2>/dev/null
  ncu -i gpurun_out/${R}_prof_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/_src.csv 2>/dev/null
  python tools/ncu_hot.py gpurun_out/_src.csv 25 > gpurun_out/${R}_$1_ncu_hot_sass.txt
  rm -f gpurun_out/${R}_prof_$1.ncu-rep gpurun_out/_src.csv
done
