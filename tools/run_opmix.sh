#!/bin/bash
# Dev tool (GPU box): one ncu --set full capture of a named shape, summarised as text under gpurun_out/.
# usage: bash tools/run_opmix.sh <tag> <shape> <n_solves> <kernel regex> <warp-steps>
TAG=$1; SHAPE=$2; N=$3; KRE=$4; WS=$5
python tools/ncu_shapes.py $SHAPE $N > gpurun_out/${TAG}_plain_$SHAPE.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$KRE" -s 1 -c 1 \
    -o gpurun_out/${TAG}_prof_$SHAPE -f python tools/ncu_shapes.py $SHAPE $N > gpurun_out/${TAG}_ncu_$SHAPE.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_prof_$SHAPE.ncu-rep gpurun_out/${TAG}_${SHAPE}_ncu_full_summary.txt $N > /dev/null
python tools/ncu_opmix.py gpurun_out/${TAG}_prof_$SHAPE.ncu-rep $WS > gpurun_out/${TAG}_${SHAPE}_ncu_opmix.txt 2>&1
ncu -i gpurun_out/${TAG}_prof_$SHAPE.ncu-rep --page source --csv --print-source sass > gpurun_out/_src.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/_src.csv 25 > gpurun_out/${TAG}_${SHAPE}_ncu_hot_sass.txt
rm -f gpurun_out/${TAG}_prof_$SHAPE.ncu-rep gpurun_out/_src.csv
