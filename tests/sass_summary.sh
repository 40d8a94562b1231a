#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md) in the built library,
# plus the MMA issue loop of both GEMM kernels.  Usage: tests/sass_summary.sh > profiles/rNN_sass_gemm.txt
so=multiplanarunet_b200/libmpunet_b200.so
echo "# cuobjdump -sass $so  (sm_100a) - mnemonic counts per kernel"
cuobjdump -sass $so | awk '
/Function :/ {f=$3; next}
{ for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTMALDG|UTMAPF|UTCBAR|UTCATOMSWS|LDTM|STTM|SYNCS|ELECT|REDG|UBLKCP|UTMASTG)/) { split($i,a,"."); k=a[1]; if ($i ~ /MULTICAST/) k=k".MULTICAST"; c[f" "k]++ } }
END { for (k in c) print k, c[k] }' | sort | awk '{k[$1]=k[$1]" "$2"="$3} END{for (f in k) print f":"k[f]}' | sort | c++filt 2>/dev/null | grep -E "mtgemm|fusion|map_fuse|sample_planes" 
for k in mtgemm_fwd_kernelILb0ELb0 mtgemm_wgrad_kernel; do
  echo
  echo "# MMA issue loop of $k (instructions between the first and the last UTCHMMA / UTCBAR of the role)"
  cuobjdump -sass $so | awk -v k="$k" '
  /Function :/ {on = index($0, k) > 0; next}
  on {line[++n]=$0; if ($0 ~ /UTCHMMA|UTCBAR/) { if (!first) first=n; last=n } }
  END { if (last - first > 260) last = first + 260; for (i=first-12; i<=last+4; i++) if (i>0) print line[i] }' | grep -v '^\s*/\* 0x' | sed -e 's/^\s*//' -e 's/ *\/\* 0x[0-9a-f]* \*\/$//' 
done
