#!/bin/bash
# SASS evidence for the claims in DESIGN.md: counts of the tell-tale mnemonics per kernel of the
# shipped library.  usage: bash tools/sass_evidence.sh > profiles/r2/sass_grep.txt
LIB=${1:-aitk.robots_b200/lib/libbrs.so}
echo "# cuobjdump -sass $LIB  (sm_100a cubin), mnemonic counts per kernel"
echo "# UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier, STG.E.EF.128 = 16-byte streaming store,"
echo "# FFMA2/FMUL2 = packed 2 x FP32, SHFL/VOTE/REDUX = warp shuffles and ballots, BAR = CTA / named barriers,"
echo "# DFMA/DMUL/DADD = float64 refinement and motion, HMMA/UTC = tensor cores (expected: 0)"
cuobjdump -sass $LIB | awk '
/Function :/ { fn=$3; next }
/^[ \t]+\/\*[0-9a-f]+\*\// {
  n[fn]++
  if ($0 ~ /UBLKCP/) a[fn,"UBLKCP"]++
  if ($0 ~ /SYNCS/) a[fn,"SYNCS"]++
  if ($0 ~ /STG\.E\.EF\.128/) a[fn,"STG.E.EF.128"]++
  if ($0 ~ /FFMA2|FMUL2/) a[fn,"FFMA2/FMUL2"]++
  if ($0 ~ /[^D]FFMA |[^D]FMUL |FADD /) a[fn,"FFMA/FMUL/FADD"]++
  if ($0 ~ /SHFL/) a[fn,"SHFL"]++
  if ($0 ~ /VOTE/) a[fn,"VOTE"]++
  if ($0 ~ /BAR\.SYNC|BAR\.ARV/) a[fn,"BAR"]++
  if ($0 ~ /DFMA|DMUL|DADD/) a[fn,"DFMA/DMUL/DADD"]++
  if ($0 ~ /MUFU/) a[fn,"MUFU"]++
  if ($0 ~ /HMMA|UTCHMMA|UTCIMMA|UTCQMMA|IMMA/) a[fn,"HMMA/UTC*MMA"]++
  if ($0 ~ /LDL|STL/) a[fn,"LDL/STL"]++
}
END {
  split("UBLKCP SYNCS STG.E.EF.128 FFMA2/FMUL2 FFMA/FMUL/FADD SHFL VOTE BAR DFMA/DMUL/DADD MUFU LDL/STL HMMA/UTC*MMA", k, " ")
  for (f in n) {
    printf "%s  instructions=%d", f, n[f]
    for (i = 1; i <= 12; i++) printf "  %s=%d", k[i], a[f,k[i]]+0
    printf "\n"
  }
}' | c++filt | sort
