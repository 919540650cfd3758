#!/usr/bin/env bash
# round 2, call AP: A/B of the two attention-backward changes (deferred flush, tail dQ on the tail warps), same box
set -u
for b in build/attn_time_v build/attn_time_vdoctic_bwd_no_defer build/attn_time_vdoctic_bwd_no_taildq build/attn_time_vdoctic_bwd_no_deferdoctic_bwd_no_taildq; do
  echo "== $b"
  for o in 1 0; do timeout 60 $b 128 b $o 1 | head -1; done
  timeout 60 $b 128 b 1 1 | head -1
done
