#!/bin/bash
# SASS evidence of the final binary: per kernel, how many tcgen05 / TMA / TMEM instructions it contains (static counts from
# cuobjdump -sass; mnemonics per /opt/skills/guides/B200_PROFILING.md).  Usage: tools/sass_digest.sh > profiles/r02_sass_digest.txt
LIB=${1:-open_clip_b200/libclipn.so}
echo "# cuobjdump -sass $LIB (sm_100a) — static instruction counts per kernel"
echo "# UTCHMMA = tcgen05.mma kind::f16, UTMALDG/UTMASTG/UTMAREDG = TMA load / store / reduce-add, UTCBAR = tcgen05.commit,"
echo "# LDTM = tcgen05.ld, HMMA = legacy mma.sync (round-1 attention kernels kept behind CLIPN_ATTN_TC=0), F2 = FFMA2/FMUL2/FADD2"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function : /{fn=$3; seen[fn]=1}
/ UTCHMMA|UTCQMMA|UTCIMMA/{mma[fn]++}
/UTMALDG/{ldg[fn]++}
/UTMASTG/{stg[fn]++}
/UTMAREDG/{red[fn]++}
/UTCBAR/{bar[fn]++}
/LDTM/{ldtm[fn]++}
/[ \t]HMMA/{hmma[fn]++}
/MUFU\.EX2/{ex2[fn]++}
/FFMA2|FMUL2|FADD2/{f2[fn]++}
END{for (f in seen) printf "%s UTCHMMA=%d UTMALDG=%d UTMASTG=%d UTMAREDG=%d UTCBAR=%d LDTM=%d HMMA=%d MUFU.EX2=%d F2=%d\n", f, mma[f], ldg[f], stg[f], red[f], bar[f], ldtm[f], hmma[f], ex2[f], f2[f]}' | c++filt | sed 's/clipn:://g' | sort
