#!/bin/bash
# ncu launch list of ONE training step (the third: plans built, side launches active, eager) of a BASELINE config with device
# time and DRAM bytes per launch -> gpurun_out/launches_<tag>_<cfg>.csv (+ profiles/dram_<cfg>.json via scripts/ncu_dram_summary.py)
# usage (on the GPU box, one GPU): scripts/ncu_list.sh <cfg> <tag>
CFG=${1:-c2}; TAG=${2:-r2}
python scripts/prof_step.py $CFG 3 > gpurun_out/plain_${TAG}_${CFG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}_${CFG}.log; exit 1; }
PER_STEP=$(grep -o "launches/step [0-9]*" gpurun_out/plain_${TAG}_${CFG}.log | grep -o "[0-9]*$")
# setup = 2 launches (one-launch weight re-pack + skip-bias table), then PER_STEP per step
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip $((2 + 2 * PER_STEP)) -c $PER_STEP --csv \
    --log-file gpurun_out/launches_${TAG}_${CFG}.csv python scripts/prof_step.py $CFG 3 > gpurun_out/ncu_${TAG}_${CFG}_list.log 2>&1
echo "per_step $PER_STEP"; tail -n 2 gpurun_out/ncu_${TAG}_${CFG}_list.log
python scripts/launch_summary.py gpurun_out/launches_${TAG}_${CFG}.csv > gpurun_out/launches_${TAG}_${CFG}_summary.txt 2>&1; head -12 gpurun_out/launches_${TAG}_${CFG}_summary.txt
