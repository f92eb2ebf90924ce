#!/bin/bash
# time a list of library builds (variants/*.so) on a list of (deck, packets, mode) cases with one forced kernel variant
# usage: tools/ab_libs.sh VARIANT "lib1 lib2 ..." "deck:n:mode ..."
V=$1; LIBS=$2; CASES=$3
for lib in $LIBS; do
  for cs in $CASES; do
    IFS=: read deck n mode <<< "$cs"
    echo -n "$(basename $lib) v$V "; SMCRT_LIB=$PWD/$lib SMCRT_VARIANT_FORCE=$V timeout 120 python tools/steady_time.py $deck $n $mode 2>&1 | tail -1 | cut -c1-120
  done
done
