#!/bin/bash
# Evidence that the shipped kernels are Blackwell-native: SASS opcode histogram + excerpts of the default rollout kernel
# (tcgen05 = UTCHMMA, TMEM ld/st = LDTM/STTM, bulk-copy engine = UBLKCP, commit = UTCBAR) and of the lean binning kernel.
#   scripts/sass_evidence.sh > profiles/r02_sass_evidence.txt
cd "$(dirname "$0")/.."
SO=density_planner_b200/libdensity_planner_b200.so
K=$(cuobjdump -sass $SO | grep "Function :" | grep "rollout_tc_kernelILi1ELi1ELi0ELi0ELi1ELi0E" | sed 's/.*Function : //')
echo "== $SO, default rollout kernel: $K  (TP=1, K-chunked, single reference, sign-free layer 2, V=0)"
cuobjdump -sass -fun "$K" $SO 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/k1.sass
echo "static instructions: $(wc -l < /tmp/k1.sass)"
echo "-- opcode histogram (top 40)"
awk '{print $2}' /tmp/k1.sass | sed 's/^@!\?U\?P[0-9T]*$//' | awk 'NF' > /tmp/k1.ops
awk '{ if ($2 ~ /^@/) print $3; else print $2 }' /tmp/k1.sass | sed 's/;$//' | sort | uniq -c | sort -rn | head -40
echo "-- Blackwell-specific opcodes"
for op in UTCHMMA UTCBAR LDTM STTM UBLKCP SYNCS LDGSTS FFMA2 FMUL2 FADD2 FHADD ELECT REDUX; do
  printf "%-8s %d\n" $op $(grep -c "\b$op" /tmp/k1.sass)
done
echo "-- excerpt: bulk copy of the operand image (UBLKCP) and TMEM allocation"
grep -n -m1 "UBLKCP" /tmp/k1.sass | cut -d: -f1 | { read n; sed -n "$((n-4)),$((n+6))p" /tmp/k1.sass; } | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-120
echo "-- excerpt: layer-1 epilogue -> STTM, K-chunk barrier, elected thread issues UTCHMMA, UTCBAR commit"
grep -n -m1 "UTCHMMA" /tmp/k1.sass | cut -d: -f1 | { read n; sed -n "$((n-22)),$((n+18))p" /tmp/k1.sass; } | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-120
echo "-- excerpt: accumulator drain (LDTM) and in-place operand store (STTM) of the layer-2 epilogue"
grep -n "LDTM.x16" /tmp/k1.sass | sed -n 5p | cut -d: -f1 | { read n; sed -n "$((n-2)),$((n+14))p" /tmp/k1.sass; } | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-120
KB=$(cuobjdump -sass $SO | grep "Function :" | grep "bin2d_lean_kernelILb0ELb0ELb1ELi768E" | sed 's/.*Function : //')
echo
echo "== lean binning kernel: $KB"
cuobjdump -sass -fun "$KB" $SO 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/k3.sass
echo "static instructions: $(wc -l < /tmp/k3.sass); ATOMS $(grep -c ATOMS /tmp/k3.sass), REDG $(grep -c REDG /tmp/k3.sass), LDG.E.128 $(grep -c 'LDG.E.128' /tmp/k3.sass), MATCH $(grep -c MATCH /tmp/k3.sass), F2I $(grep -c F2I /tmp/k3.sass)"
