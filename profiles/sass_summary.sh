#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): run on the CPU box.
# usage: profiles/sass_summary.sh > profiles/rNN_sass_summary.txt
so=${1:-vox_serve_b200/lib/libvoxb200.so}
echo "# cuobjdump -sass $so | per-kernel mnemonic counts  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (linear TMA), UTMALDG = cp.async.bulk.tensor, HMMA = mma.sync,"
echo "# CCTL.PF* = prefetch.global.L2 (L2 prefetcher), REDG/ATOMG = progress words / split-KV arrival counters"
cuobjdump -sass "$so" 2>/dev/null | awk '
/Function :/ {fn=$3}
/UTC[A-Z]*MMA/ {mma[fn]++; next}
/UBLKCP/ {blk[fn]++}
/UTMALDG/ {tma[fn]++}
/LDTM/ {ldtm[fn]++}
/[^A-Z]HMMA/ {hmma[fn]++}
/CCTL[A-Z.]*PF/ {pf[fn]++}
END {
  printf "%-78s %8s %7s %8s %6s %6s %8s\n", "kernel", "UTC*MMA", "UBLKCP", "UTMALDG", "LDTM", "HMMA", "CCTL.PF";
  for (f in mma) seen[f]=1; for (f in blk) seen[f]=1; for (f in tma) seen[f]=1; for (f in ldtm) seen[f]=1; for (f in hmma) seen[f]=1; for (f in pf) seen[f]=1;
  for (f in seen) printf "%-78s %8d %7d %8d %6d %6d %8d\n", substr(f,1,78), mma[f], blk[f], tma[f], ldtm[f], hmma[f], pf[f] | "sort";
}'
