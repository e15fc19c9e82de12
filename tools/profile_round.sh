#!/bin/bash
# Round profile set (B200_PROFILING.md recipe): launch lists of the headline and the short-block workloads, and one
# ncu --set full capture of the dominant kernel of each workload family.  Outputs under gpurun_out/<tag>_*.
TAG="${1:-r02}"
Q="--quick --steps 2 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_genomic.csv python bench.py --workload genomic1 $Q > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_short.csv python bench.py --workload short2k --evolve $Q > /dev/null 2>&1
cap() {  # name workload kernel-regex extra
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s 2 -c 1 -o gpurun_out/${TAG}_$1 python bench.py --workload $2 $4 $Q > /dev/null 2>&1
}
cap k_dp_reg genomic1 'k_dp_reg' ''
cap k_dp_smpf short2k 'k_dp_smpf' '--evolve'
cap k_dp_smp_chunked hundred_short 'k_dp_smp<' '--evolve'
cap k_dp_chain hundred 'k_dp_chain' '--evolve'
cap k_pack short2k 'k_pack\(' '--evolve'
cap k_pack2 short2k 'k_pack2' '--evolve'
cap k_evolve short2k 'k_evolve' '--evolve'
cap k_dp_smps mid5 'k_dp_smps' '--evolve'
ls -la gpurun_out/${TAG}_*
