#!/bin/bash
# Round profile set (B200_PROFILING.md recipe), run on the GPU box: launch lists of the headline and the short-block steps and
# one ncu --set full capture of the dominant kernel of each workload family.  The captures are summarised on the box
# (tools/profile_summary.py, tools/src_regions.py) and only the summaries (profiles/<tag>_*.md + <tag>_traffic.json) come back
# through gpurun_out/ -- the .ncu-rep files are 20 MB each.
TAG="${1:-r02}"
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
Q="--quick --steps 2 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_genomic.csv python bench.py --workload genomic1 $Q > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_short.csv python bench.py --workload short2k --evolve $Q > /dev/null 2>&1
python tools/profile_summary.py ${TAG}_genomic $OUT/launches_genomic.csv /nonexistent
python tools/profile_summary.py ${TAG}_short $OUT/launches_short.csv /nonexistent
echo "{" > $OUT/traffic.json
cap() {  # label workload kernel-regex extra-bench-args
  rm -f /tmp/cap.ncu-rep
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s 2 -c 1 -o /tmp/cap python bench.py --workload $2 $4 $Q > /dev/null 2>&1
  if [ -f /tmp/cap.ncu-rep ]; then
    python tools/profile_summary.py $TAG /nonexistent /tmp/cap.ncu-rep $1
    ncu -i /tmp/cap.ncu-rep --page source --csv > /tmp/cap_src.csv 2>/dev/null
    python tools/src_regions.py /tmp/cap_src.csv 1.0 > $OUT/${1}_regions.txt 2>&1
    ncu -i /tmp/cap.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[2]))
def b(k):
    v=float(d[k]); u=dict(zip(rows[0],rows[1]))[k]
    return v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
tu=dict(zip(rows[0],rows[1]))['gpu__time_duration.sum']
ms=float(d['gpu__time_duration.sum'])*{'ns':1e-6,'us':1e-3,'usecond':1e-3,'ms':1.0,'msecond':1.0,'s':1e3,'second':1e3,'nsecond':1e-6}.get(tu,1.0)
print('  \"$1\": {\"workload\": \"$2\", \"dram_bytes_per_launch\": %d, \"ms\": %.6f},'%(b('dram__bytes_read.sum')+b('dram__bytes_write.sum'), ms))" >> $OUT/traffic.json
  else
    echo "capture of $1 failed" >> $OUT/errors.txt
  fi
}
cap k_dp_reg genomic1 '^k_dp_reg$' ''
cap k_dp_smpf short2k '^k_dp_smpf$' '--evolve'
cap k_dp_smp_chunked hundred_short '^k_dp_smp$' '--evolve'
cap k_dp_chain hundred '^k_dp_chain$' '--evolve'
cap k_dp_smps mid5 '^k_dp_smps$' '--evolve'
cap k_sigma_p2 hundred_short '^k_sigma_p2$' '--evolve'
cap k_sigma_rows3 hundred '^k_sigma_rows3$' '--evolve'
cap k_dp_regtu genomic_lowgap '^k_dp_regtu$' ''
cap k_pack short2k '^k_pack$' '--evolve'
cap k_pack2 short2k '^k_pack2$' '--evolve'
cap k_evolve short2k '^k_evolve$' '--evolve'
cap k_hss_thr short2k '^k_hss_thr$' '--evolve'
echo '  "_": null' >> $OUT/traffic.json; echo "}" >> $OUT/traffic.json
cp profiles/${TAG}_*.md $OUT/ 2>/dev/null
ls -la $OUT
