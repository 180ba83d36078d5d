#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / bulk-copy use (B200_PROFILING.md):
#   UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,
#   UTMALDG / UTMASTG = tensor-map TMA loads / stores (none here: operands are fp32 in HBM and are split into bf16x3
#   planes by the loader warps, weights are pre-packed contiguous blocks that one bulk copy moves).
# usage: scripts/sass_summary.sh > profiles/sass_r2.txt
cd "$(dirname "$0")/.."
cuobjdump -sass neuralsat_b200/libcrown_b200.so | awk '
  /Function :/ { name=$3; if (!(name in seen)) { order[++n]=name; seen[name]=1 } }
  /UTCHMMA/ {mma[name]++} /LDTM/ {ld[name]++} /UTCBAR/ {bar[name]++} /UBLKCP/ {blk[name]++} /UTMALDG|UTMASTG/ {tma[name]++}
  /FFMA/ {ffma[name]++}
  END { printf "%-90s %8s %6s %7s %7s %8s %7s\n", "kernel", "UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMA*", "FFMA";
        for (i=1;i<=n;i++) { k=order[i]; if (mma[k]+ld[k]+bar[k]+blk[k]+tma[k] > 0)
          printf "%-90s %8d %6d %7d %7d %8d %7d\n", substr(k,1,90), mma[k], ld[k], bar[k], blk[k], tma[k], ffma[k] } }'
