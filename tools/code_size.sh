#!/bin/bash
# Dev tool: static SASS size of one kernel instance by source region (the instruction cache is the budget to watch).
#   tools/code_size.sh lean_16 "fdm_lean_kernelILi16ELb1ELb0" [top]
set -e
obj=${1:-lean_16}
pat=${2:-fdm_lean_kernelILi16ELb1ELb0}
top=${3:-20}
root=$(dirname "$(dirname "$(readlink -f "$0")")")
work=$root/build/elf_$obj
mkdir -p "$work"
( cd "$work" && cuobjdump -xelf all "$root/build/obj/default/$obj.o" > /dev/null 2>&1 && nvdisasm --print-line-info ./*.cubin > dis.txt 2> /dev/null )
python "$root/tools/sass_lines.py" "$work/dis.txt" "$pat" "$top"
