#!/usr/bin/env bash
# round 2, call AE: staged dQ, final form: kernel timing + in-step event tables, staged vs two-pass on the same box
set -u
mkdir -p gpurun_out
for o in 1 0; do for st in 1 0; do timeout 60 build/attn_time 128 b $o $st | head -1; done; done
for st in 1 0; do
  OCTIC_ATTN_STAGED_DQ=$st timeout 200 python tools/profile_step.py --batch 128 --events > gpurun_out/r2ae_events_b128_staged$st.txt 2>&1; echo "events staged=$st rc=$?"
  grep -E "step total|attention" gpurun_out/r2ae_events_b128_staged$st.txt
done
