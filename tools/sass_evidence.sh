#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove Blackwell-native code paths
# (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG).
#   tools/sass_evidence.sh > profiles/sass_opcodes.txt
set -e
cd "$(dirname "$0")/.."
so=sbmc_b200/libsbmc_b200.so
echo "# cuobjdump -sass $so  ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# kernel | UTMALDG | UTMASTG | UTCHMMA | LDTM | STTM | UTCBAR(commit) | SYNCS(mbarrier) | HMMA(legacy)"
cuobjdump -sass "$so" | awk '
  /Function :/ { if (name != "") print name " | " a " | " b " | " c " | " d " | " e " | " f " | " g " | " h;
                 name=$3; a=b=c=d=e=f=g=h=0 }
  /UTMALDG/ {a++} /UTMASTG/ {b++} /UTCHMMA/ {c++} /LDTM/ {d++} /STTM/ {e++} /UTCBAR/ {f++} /SYNCS/ {g++} / HMMA/ {h++}
  END { print name " | " a " | " b " | " c " | " d " | " e " | " f " | " g " | " h }' | while IFS= read -r line; do
    mangled=${line%% |*}
    rest=${line#* |}
    echo "$(echo "$mangled" | c++filt | cut -c1-110) |$rest"
  done | sort
