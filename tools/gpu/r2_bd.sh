#!/usr/bin/env bash
# round 2, call BD: life times of the attention-backward CTAs per SM (busy time vs gaps) at the full batch
set -u
timeout 60 build/attn_trace 128 b 1 1 | grep -E "per launch|CTA life"
timeout 60 build/attn_trace 128 b 0 1 | grep -E "per launch|CTA life"
