#!/usr/bin/env bash
# round 2, call L: attention backward, round-1 kernel vs current, same binary harness, batch 128
set -u
for v in attn_time_old attn_time; do
  for o in 1 0; do echo "$v: $(timeout 60 build/$v 128 b $o | head -1)"; done
done
