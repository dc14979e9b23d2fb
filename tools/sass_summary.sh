#!/bin/bash
# Opcode evidence for the tcgen05 / TMA / TMEM / FP64-tensor paths in the shipped library (runs without a GPU).
SO=pybo_b200/lib/libbo_b200.so
OUT=${1:-profiles/r2_sass_summary.txt}
cuobjdump -sass $SO > /tmp/bo_sass.txt
{
  echo "# cuobjdump -sass $SO | grep -c <opcode>   (sm_100a; $(stat -c %s $SO) bytes; $(date -u +%Y-%m-%dT%H:%MZ))"
  for op in UTCIMMA UTCBAR UTCCP LDTM STTM UTMALDG UBLKCP UTMAPF SYNCS DMMA DFMA MUFU.RCP64H MUFU.RSQ64H LDGSTS ATOMG "LD.E.*STRONG.GPU" "ST.E.*STRONG.GPU" MEMBAR; do
    printf "%-22s %s\n" "$op" "$(grep -c -E "$op" /tmp/bo_sass.txt)"
  done
  echo
  echo "# per kernel (function name -> UTCIMMA / LDTM / UTMALDG / UBLKCP / DMMA counts), kernels with any of them"
  awk '/Function :/ {name=$3} /UTCIMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /UBLKCP/ {d[name]++} /DMMA/ {e[name]++} END {for (n in a) seen[n]=1; for (n in b) seen[n]=1; for (n in c) seen[n]=1; for (n in d) seen[n]=1; for (n in e) seen[n]=1; for (n in seen) printf "%s %d %d %d %d %d\n", n, a[n], b[n], c[n], d[n], e[n]}' /tmp/bo_sass.txt | c++filt | sort | awk '{cnt=$(NF-4)" "$(NF-3)" "$(NF-2)" "$(NF-1)" "$NF; $NF="";$(NF-1)="";$(NF-2)="";$(NF-3)="";$(NF-4)=""; n=$0; sub(/\(.*/,"",n); printf "%-70s %s\n", n, cnt}' | sort | uniq
} > $OUT
head -30 $OUT
