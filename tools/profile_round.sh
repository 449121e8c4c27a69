#!/bin/bash
# tools/profile_round.sh — ncu evidence for one AES round (all kernels of the hot path), run on the GPU box:
#   1. launch list with per-launch durations (gpu__time_duration.sum, --clock-control none)
#   2. ncu --set full capture of the 9 consecutive launches that make up one round, located from the launch list
# Outputs under gpurun_out/: launches_$TAG.csv, prof_round_$TAG.ncu-rep
TAG=${1:-r1}
BLOCKS=${2:-48}
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 1 --warmup 1 --blocks $BLOCKS --no-cpu-baseline --no-key-expansion"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
SKIP=$(python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name")
names=[r[ki] for r in rows[1:]]
# each launch appears once per metric; here one metric -> one row per launch
idx=[i for i,n in enumerate(names) if "umma_digit_tiles_kernel" in n]
# rounds start at every second digit-tiles launch (ks digits, then pfks digits); take the 3rd round of the run
starts=idx[0::2]
print(starts[2] if len(starts)>2 else starts[0])
PY
)
echo "round starts at launch $SKIP" >> gpurun_out/ncu_launches_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip $SKIP --launch-count 9 -f -o gpurun_out/prof_round_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
# condensed evidence next to the report; KEEP_REP=0 drops the (≈ 40 MB) report itself — gpurun copies back at most 64 MiB
python tools/ncu_summary.py gpurun_out/prof_round_$TAG.ncu-rep gpurun_out/round_$TAG > /dev/null
# PHASE_KERNEL=regex: stall samples per barrier-delimited phase of that kernel (tools/ncu_phase_table.py)
if [ -n "$PHASE_KERNEL" ]; then python tools/ncu_phase_table.py gpurun_out/prof_round_$TAG.ncu-rep "$PHASE_KERNEL" > gpurun_out/phases_$TAG.md 2>&1; fi
if [ "${KEEP_REP:-1}" = "0" ]; then rm -f gpurun_out/prof_round_$TAG.ncu-rep; fi
