#!/usr/bin/env bash
# Evidence that the GEMM kernels are tcgen05 / TMEM / TMA code and a register / shared-memory table of every kernel:
#   profiles/r2_sass_opcodes.txt   per-kernel counts of UTCHMMA / UTCHMMA.2CTA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA
#                                  loads), UTCBAR (tcgen05.commit), SYNCS (mbarrier) in the SASS of the in-tree libraries
#   profiles/r2_ptxas_resources.txt  registers / shared memory / spills per kernel (cuobjdump --dump-resource-usage)
# Runs without a GPU (cuobjdump reads the cubin inside the .so).
set -euo pipefail
cd "$(dirname "$0")/.."
OUT=profiles/r2_sass_opcodes.txt
: > $OUT
for lib in simulgen_vae_b200/libsimulgen_b200.so simulgen_vae_b200/libsimulgen_b200_fp16.so; do
  echo "== $lib ($(cuobjdump -lelf $lib | head -1))" >> $OUT
  printf "%-12s %-8s %-6s %-8s %-7s %-6s %s\n" "UTCHMMA.2CTA" "UTCHMMA" "LDTM" "UTMALDG" "UTCBAR" "SYNCS" "kernel" >> $OUT
  cuobjdump -sass $lib | awk '
    /Function :/ { fn=$3 }
    /UTCHMMA\.2CTA/ { a[fn]++ ; next }
    /UTCHMMA/ { b[fn]++ }
    /LDTM/ { c[fn]++ }
    /UTMALDG/ { d[fn]++ }
    /UTCBAR/ { e[fn]++ }
    /SYNCS/ { f[fn]++ }
    END { for (k in d) printf "%-12d %-8d %-6d %-8d %-7d %-6d %s\n", a[k], b[k], c[k], d[k], e[k], f[k], k }' | sort -k7 >> $OUT
done
echo "(UTCHMMA = tcgen05.mma kind::f16, .2CTA = cta_group::2; LDTM = tcgen05.ld; UTMALDG = cp.async.bulk.tensor; UTCBAR = tcgen05.commit)" >> $OUT
RES=profiles/r2_ptxas_resources.txt
: > $RES
for lib in simulgen_vae_b200/libsimulgen_b200_fp16.so; do
  echo "== $lib" >> $RES
  cuobjdump --dump-resource-usage $lib 2>/dev/null | grep -A1 "Function" | grep -o "Function [^:]*\|REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*" | paste - - - - | \
    awk '{gsub("Function ","",$0); print $2, $3, $4, $1}' | while read r s sh fn; do echo "$r $s $sh $(echo $fn | c++filt | cut -c1-150)"; done | sort -k4 >> $RES
done
wc -l $OUT $RES
