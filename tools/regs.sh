#!/bin/bash
# usage: tools/regs.sh <lib.so>  -- per-kernel registers / stack (spill) from the cubin
cuobjdump --dump-resource-usage "$1" 2>/dev/null | grep -E "Function|REG" | paste - - | \
  sed -E 's/.*fdm_warp_kernelILi([0-9]+)ELb([01])ELb([01])[^ ]* +REG:([0-9]+) STACK:([0-9]+).*/C=\1 AMER=\2 TRACE=\3 REG=\4 STACK=\5/' | grep -v fp64
