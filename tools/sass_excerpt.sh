#!/bin/bash
# profiles/sass_<tag>.txt: what the built library holds -- target arch, per-kernel registers / stack / shared memory
# (cuobjdump --dump-resource-usage), and the SASS mnemonics that show the Blackwell data-movement paths in use
# (UBLKCP = TMA bulk copy, LDG.E.*.256 = 256-bit loads, ATOMS = shared-memory atomics), per kernel.
TAG=${1:-r02}; LIB=graspit_b200/libb2c.so; OUT=profiles/sass_$TAG.txt
{
  echo "# $(date -u +%F) $LIB  ($(stat -c %s $LIB) bytes), built by python -m graspit_b200.build (nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
  echo "## ELF targets"; cuobjdump -lelf $LIB
  echo; echo "## resource usage per kernel (REG, STACK, SHARED, LOCAL; cuobjdump --dump-resource-usage)"
  cuobjdump --dump-resource-usage $LIB | c++filt | grep -A1 "Function " | grep -v "^--" | paste - - | sed 's/Function \(.*\):\s*/\1 | /' | cut -c1-260
  echo; echo "## SASS mnemonic counts per kernel"
  cuobjdump -sass $LIB | c++filt | awk '
    /Function :/ {name=$3; sub(/\(.*/,"",name)}
    /UBLKCP/ {u[name]++} /LDG\.E\.[A-Z0-9.]*256/ {l[name]++} /ATOMS/ {a[name]++} /LDL|STL/ {s[name]++} /SYNCS|MBAR/ {m[name]++} /REDUX|VOTE/ {v[name]++} /DFMA|DADD|DMUL/ {d[name]++}
    {n[name]++}
    END {printf "%-44s %8s %8s %8s %8s %8s %8s %8s %8s\n","kernel","instr","UBLKCP","LDG.256","mbarrier","ATOMS","VOTE","fp64","LDL/STL"; for (k in n) if (k!="") printf "%-44s %8d %8d %8d %8d %8d %8d %8d %8d\n",k,n[k],u[k],l[k],m[k],a[k],v[k],d[k],s[k]}' | sort
  echo; echo "## excerpt: TMA bulk copy of the tree top in energy_kernel, and the 256-bit sibling-pair load"
  cuobjdump -sass $LIB | c++filt | awk '/Function : .*energy_kernel\(/ {f=1} /Function :/ && !/energy_kernel\(/ {f=0} f' | grep -E "UBLKCP|SYNCS|LDG\.E\.[A-Z0-9.]*256|PRMT .*0x7610|PRMT .*0x7632" | head -24
} > $OUT
wc -l $OUT
