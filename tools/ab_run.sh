#!/bin/bash
# tools/ab_run.sh WORKLOAD NAME... : runs bench.py once per library variant and prints the stage times
cd "$(dirname "$0")/.."
w=$1; shift
for n in "$@"; do
  RNACODE_CUDA_LIB=$PWD/rnacode_b200/lib/ab/libRNAcode_cuda_$n.so python bench.py --no-cpu --workload $w --steps 3 > /tmp/ab_$n.json 2> /tmp/ab_$n.err
  python -c "
import json; d=json.loads(open('/tmp/ab_$n.json').read().strip().splitlines()[-1]); print('$w %-10s'%'$n', '%.3e'%d['value'], 'dp %.3f ms'%d['roofline']['stage_ms_per_step']['dp'])" 2>/dev/null || tail -3 /tmp/ab_$n.err
done
