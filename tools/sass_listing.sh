#!/usr/bin/env bash
# Per-kernel count of the Blackwell-native SASS mnemonics in the built library (developer tool):
# UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add,
# LDTM / STTM = tcgen05.ld / st (TMEM), HMMA = legacy mma.sync (must be 0).
# usage: tools/sass_listing.sh > profiles/rN_sass_listing.txt
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SO="${1:-${HERE}/../omnihuman-1-hack_b200/libb200dit.so}"
echo "# cuobjdump -sass $(basename "$SO") | per-function mnemonic counts (kernels that use the tensor cores or TMA)"
cuobjdump -sass "$SO" 2>/dev/null | awk '
/Function :/ {fn=$3}
/UTCHMMA/ {m[fn]++; if ($0 ~ /2CTA/) m2[fn]++; any[fn]=1}
/UTMALDG/ {l[fn]++; any[fn]=1}
/UTMASTG/ {st[fn]++; any[fn]=1}
/UTMAREDG/ {rd[fn]++; any[fn]=1}
/LDTM/ {lt[fn]++; any[fn]=1}
/STTM/ {stm[fn]++; any[fn]=1}
/ HMMA/ {h[fn]++; any[fn]=1}
END {for (f in any) printf "%s UTCHMMA=%d (2CTA %d) UTMALDG=%d UTMASTG=%d UTMAREDG=%d LDTM=%d STTM=%d HMMA=%d\n", f, m[f]+0, m2[f]+0, l[f]+0, st[f]+0, rd[f]+0, lt[f]+0, stm[f]+0, h[f]+0}' | c++filt | sed -e 's/CUtensorMap_st/CUtensorMap/g' | sort
echo "# totals"
cuobjdump -sass "$SO" 2>/dev/null | awk '/UTCHMMA/{a++} /UTCHMMA.*2CTA/{b++} /UTMALDG/{c++} /UTMASTG/{d++} /UTMAREDG/{e++} /LDTM/{f++} /STTM/{g++} / HMMA/{h++} END{printf "UTCHMMA=%d (2CTA %d) UTMALDG=%d UTMASTG=%d UTMAREDG=%d LDTM=%d STTM=%d HMMA=%d\n",a,b,c,d,e,f,g,h+0}'
