#!/bin/bash
# tools/cap_one.sh LABEL WORKLOAD KERNEL-REGEX SKIP [bench args]: one ncu --set full capture of a kernel on the GPU box; the
# per-instruction source page and the raw metrics come back as CSV under gpurun_out/ (the .ncu-rep stays on the box).
L=$1; W=$2; K=$3; S=$4; shift 4
rm -f /tmp/cap.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -o /tmp/cap python bench.py --workload $W --quick --steps 2 --warmup 1 "$@" > /dev/null 2>&1
ncu -i /tmp/cap.ncu-rep --page source --csv > gpurun_out/${L}_src.csv 2>/dev/null
ncu -i /tmp/cap.ncu-rep --page raw --csv > gpurun_out/${L}_raw.csv 2>/dev/null
python tools/src_regions.py gpurun_out/${L}_src.csv 1.0
