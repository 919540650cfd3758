#!/usr/bin/env bash
# round 2, call X: per-op event table + launch list (ncu time-only) of one training step at the current state
set -u
mkdir -p gpurun_out
timeout 200 python tools/profile_step.py --batch 128 --events > gpurun_out/r2x_events_b128.txt 2>&1; echo "events rc=$?"
cat gpurun_out/r2x_events_b128.txt | tail -28
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_launches_b64.csv python tools/profile_step.py --batch 64 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2x_launches_b64.csv > gpurun_out/r2x_launches_b64.txt 2>&1; head -60 gpurun_out/r2x_launches_b64.txt
