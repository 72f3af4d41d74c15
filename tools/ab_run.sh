#!/bin/bash
# A/B runs of library variants on one GPU box (used under gpurun): per-stage CUDA-event times of a C2/C3/C4 pass.
#   tools/ab_run.sh "default r1 ss4" "c2 c3"      variants: default | NAME (build/variants/libbpt_NAME.so) | NAME@ENV=VAL,ENV=VAL
mkdir -p gpurun_out
V="${1:-default}"; CFGS="${2:-c2}"
for v in $V; do
  name="${v%%@*}"; envs=""
  if [[ "$v" == *@* ]]; then envs="${v#*@}"; envs="${envs//,/ }"; fi
  lib=""
  if [ "$name" != "default" ]; then lib="BPT_LIBRARY=$PWD/buas_pathtracer_b200/csrc/build/variants/libbpt_$name.so"; fi
  for c in $CFGS; do
    spp=""; [ "$c" = "c3" ] && spp="--spp 64"; [ "$c" = "c4" ] && spp="--spp 64"
    echo "=== $v $c detail" | tee -a gpurun_out/ab.log
    env $lib $envs timeout 300 python tools/profile_pass.py --config $c --passes 3 $spp 2>&1 | grep -v "^Constructed\|^Nodes with" | tee -a gpurun_out/ab.log
    echo "=== $v $c pipelined" | tee -a gpurun_out/ab.log
    env $lib $envs timeout 300 python tools/profile_pass.py --config $c --passes 3 $spp --no-detail 2>&1 | grep "Mrays\|total_ms" | tee -a gpurun_out/ab.log
  done
done
