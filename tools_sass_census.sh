#!/bin/bash
# profiles/sass_census.txt: the Blackwell-native instruction census of the shipped library (B200_PROFILING.md:
# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP, cp.async -> LDGSTS, legacy mma.sync -> HMMA).
set -e
cd "$(dirname "$0")"
LIB=enerf_b200/libenerf_b200.so
OUT=profiles/sass_census.txt
{
  echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | tail -1))"
  echo "# whole library"
  cuobjdump -sass "$LIB" | grep -oE "\b(UTCHMMA|UTCQMMA|UTCIMMA|LDTM|STTM|UTCBAR|UTMALDG|UTMASTG|UBLKCP|LDGSTS|HMMA|SYNCS)\b[.A-Z0-9_]*" | sed 's/\..*//' | sort | uniq -c | sort -rn
  echo "# per object"
  for o in enerf_b200/csrc/build/*.o; do
    c=$(cuobjdump -sass "$o" | grep -oE "\b(UTCHMMA|LDTM|UTMALDG|UBLKCP|LDGSTS|HMMA)\b" | sort | uniq -c | tr '\n' ' ')
    [ -n "$c" ] && echo "$(basename $o): $c"
  done
  echo "# kernels containing UTMALDG (tensor-map TMA)"
  cuobjdump -sass "$LIB" | awk '/Function :/ {f=$3} /UTMALDG/ {print f}' | sort | uniq -c | c++filt | cut -c1-160
} > "$OUT"
cat "$OUT" | head -40
