#!/bin/bash
# One gpurun call that produces every artefact profiles/ needs for a round (B200_PROFILING.md recipe):
#   tools/profile_round.sh r02a        (run from the repo root; needs ~4 GPU-minutes)
# 1. plain bench (must exit 0 before any ncu pass), 2. ncu launch list of the same command, 3. ncu --set full of the dominant
# GEMM alone (tools/prof_dominant.py) and of the ingest kernel, 4. summaries written to gpurun_out/ for copying into profiles/.
set -u
TAG=${1:?usage: tools/profile_round.sh <tag, e.g. r02a>}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
/usr/local/graft/bin/gpurun --timeout 420 -- "
  timeout 120 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
  timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
  timeout 60 python tools/prof_dominant.py > gpurun_out/${TAG}_dom_plain.log 2>&1 &&
  timeout 90 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 4 -c 1 -o gpurun_out/${TAG}_dominant_gemm python tools/prof_dominant.py > gpurun_out/${TAG}_ncu_dom.log 2>&1
  tail -c 400 gpurun_out/${TAG}_plain.log
"
python tools/summarize_ncu.py launches gpurun_out/${TAG}_launches.csv > profiles/${TAG}_launches.md
python tools/summarize_ncu.py full gpurun_out/${TAG}_dominant_gemm.ncu-rep > profiles/${TAG}_dominant_gemm_full.md
echo "wrote profiles/${TAG}_launches.md profiles/${TAG}_dominant_gemm_full.md"
