#!/bin/bash
# ncu launch list of one bench workload (B200_PROFILING.md recipe): per-kernel time, cold-cache and serialised.
# usage: tools/launch_list.sh <workload> <out.csv> [extra bench args]
W="$1"; OUT="$2"; shift 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT" \
  python bench.py --workload "$W" --quick --evolve --steps 2 --warmup 1 "$@" > /dev/null 2>&1
python - "$OUT" <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    name = r[ki].split("(")[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%4d total %9.3f ms avg %8.3f ms share %5.1f%%" % (k[:60], n, ms, ms / n, 100 * ms / tot))
PY
