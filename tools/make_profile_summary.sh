#!/bin/bash
# Turn gpurun_out/<tag>_prof.ncu-rep (+ <tag>_launches.csv, <tag>_bench.json) into text summaries under profiles/.
TAG=$1
[ -f gpurun_out/${TAG}_launches.csv ] && cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
[ -f gpurun_out/${TAG}_bench.json ] && cp gpurun_out/${TAG}_bench.json profiles/${TAG}_bench.json
if [ -f gpurun_out/${TAG}_prof.ncu-rep ]; then
  ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > /tmp/${TAG}_raw.csv 2>/dev/null
  python tools/ncu_raw_summary.py /tmp/${TAG}_raw.csv > profiles/${TAG}_ncu_full_metrics.txt
  ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source sass > /tmp/${TAG}_src.csv 2>/dev/null
  python tools/ncu_src_summary.py /tmp/${TAG}_src.csv 8 | awk '/^===/{k=$0; if (seen[k]++) skip=1; else skip=0} !skip' > profiles/${TAG}_ncu_source_summary.txt
fi
ls -la profiles | grep ${TAG}
