#!/bin/bash
# Regenerates profiles/sass_summary.txt from the in-tree sm_100a libraries (no GPU needed).
cd "$(dirname "$0")/.."
out=profiles/sass_summary.txt
: > $out
for so in gemm coll ops; do
  f=batch_shipyard_b200/_native/libshipyard_$so.so
  echo "== $f ($(cuobjdump -lelf $f 2>/dev/null | grep -c sm_100a) sm_100a cubins)" >> $out
  cuobjdump -sass $f 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sort | uniq -c | sort -rn \
    | grep -E "UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTCATOMSWS|SYNCS|REDG|LDGMC|\.MC|MULTIMEM|ATOMG|UBLKCP|\.STRONG\.SYS|HMMA|DFMA" >> $out
  echo >> $out
done
echo "-- per-kernel registers / smem (cuobjdump -res-usage)" >> $out
for so in gemm coll ops; do
  cuobjdump -res-usage batch_shipyard_b200/_native/libshipyard_$so.so 2>/dev/null | grep -A1 "Function" | grep -vE "^--" | paste - - | sed -E 's/ +/ /g' | cut -c1-260 >> $out
done
