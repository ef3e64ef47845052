#!/bin/bash
# Host-side phases of rt_scene_create for the 1 M-sphere scene on the GPU box (16 cores).
set -u
mkdir -p gpurun_out
RT_B200_BVH_TIMING=1 timeout 300 python scripts/e2e_breakdown.py C4 2>&1 | tail -40 | tee gpurun_out/summary_aw.txt
