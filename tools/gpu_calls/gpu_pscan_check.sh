#!/bin/bash
# tests of the partition-major scan + bench in both scan modes + launch list of the partition mode
timeout 600 python -m pytest tests -m gpu -x -q -k "partition_major or filter_path" 2>&1 | tail -5
for m in query partition; do
  FDB_FILTER_SCAN=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-scan-large > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_$m.json"))
print("$m", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["phase_ms_per_step"].items()}, round(d["roofline"]["frac"],3), d["parity"], d["query_path"])
P
done
export FDB_FILTER_SCAN=partition
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ps.csv python tools/prof_query.py 2 > gpurun_out/ncu_ps1.log 2>&1
python tools/launch_summary.py gpurun_out/launches_ps.csv 2>/dev/null | grep -i "pscan\|pmerge\|pg_\|fselect\|launches"
if [ "$1" = "full" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pscan_kernel -c 1 -o gpurun_out/prof_pscan -f python tools/prof_query.py 1 > gpurun_out/ncu_ps2.log 2>&1; tail -2 gpurun_out/ncu_ps2.log
fi
