#!/bin/bash
# SASS opcode histogram per kernel of webaudio-modem_b200/libwam.so (cuobjdump), the evidence for which units the kernels
# use: UTMALDG / SYNCS (TMA + mbarrier), FFMA2 / FMUL2 (packed f32x2), DFMA / F2F (float64 and conversions), MUFU, POPC.
# usage: scripts/sass_histogram.sh > profiles/r02_sass_opcodes.txt
cd "$(dirname "$0")/.."
lib=webaudio-modem_b200/libwam.so
echo "# $(date -u +%F) $(nvcc --version | tail -2 | head -1)"
echo "# source hash $(cat $lib.srchash 2>/dev/null | cut -c1-16)"
cuobjdump -sass $lib 2>/dev/null | awk '
/Function :/ { fn=$3; next }
/^[ \t]+\/\*[0-9a-f]+\*\// {
  op=$2; if (op ~ /^@/) op=$3; sub(/;/, "", op); split(op, p, "."); base=p[1];
  if (base=="UTMALDG" || base=="SYNCS" || base=="MUFU" || base=="F2F" || base=="LDGSTS") base=p[1] (p[2]!="" ? "." p[2] : "");
  n[fn]++; c[fn, base]++; ops[base]=1
}
END {
  for (f in n) {
    printf "== %s  (%d instructions)\n", f, n[f];
    line="";
    m=0; for (o in ops) if ((f, o) in c) { k[++m]=o }
    # sort by count, descending
    for (i=1;i<=m;i++) for (j=i+1;j<=m;j++) if (c[f,k[j]] > c[f,k[i]]) { t=k[i]; k[i]=k[j]; k[j]=t }
    for (i=1;i<=m;i++) { line = line sprintf("%s %d  ", k[i], c[f,k[i]]); if (i%10==0) { print "   " line; line="" } }
    if (line!="") print "   " line
    delete k
  }
}' | c++filt
