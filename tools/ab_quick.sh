#!/bin/bash
# A/B of library variants (rnacode_b200/lib/ab/*.so) and run-time switches on bench workloads: one "[quick]" line each.
# usage: tools/ab_quick.sh "<workload> ..." "<label>:<ENV=VAL ...>" ...   (label "x:" = defaults)
WL="$1"; shift
for w in $WL; do
  for spec in "$@"; do
    label="${spec%%:*}"; envs="${spec#*:}"
    line=$(env $envs python bench.py --workload $w --quick --evolve --steps 5 2>&1 >/dev/null | grep '\[quick\]')
    echo "[$label] $line"
  done
done
