#!/bin/bash
# Dev tool: time (and parity-test) alternative builds of the library under build/variants/ on the GPU box.
# usage (through gpurun): bash tools/run_variants.sh "<timing scripts>" "<pytest files>" lib1 lib2 ...
TIMERS=$1; TESTS=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  export FDM_B200_LIB=$PWD/build/variants/libfdm_$v.so
  for tscript in $TIMERS; do
    echo "== $v $tscript" >> gpurun_out/variants.log
    python $tscript 2>&1 | grep -v -i "warn\|from_numpy" >> gpurun_out/variants.log
  done
  if [ -n "$TESTS" ] && [ "$v" != "base" ]; then
    echo "== $v pytest $TESTS" >> gpurun_out/variants.log
    python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/variants.log
  fi
done
cat gpurun_out/variants.log
