#!/bin/bash
# ncu --set full of all conv_gemm launches of one warm detector forward (batch 16) -> gpurun_out/<tag>_raw.csv (+ source pages
# of the launches listed in $SRC_IDS, gzip'ed).  Usage: tools/ncu_conv_all.sh <tag> [launches per forward, default 74]
TAG=${1:-conv_all}
N=${2:-74}
SRC_IDS=${SRC_IDS:-"0 3 4 13 14 27"}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip $N --launch-count $N \
  -o /tmp/$TAG python tools/prof_detector.py --reps 2 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv > gpurun_out/${TAG}_summary.csv
ncu -i /tmp/$TAG.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_split_source.py gpurun_out/${TAG}_src $SRC_IDS
gzip -f gpurun_out/${TAG}_raw.csv
ls -la /tmp/$TAG.ncu-rep gpurun_out/ | tail -20
