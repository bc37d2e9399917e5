#!/bin/bash
# Round-2 profile capture (run under gpurun, ONE GPU):  bash tools/gpu_profile.sh dhfr [apoa1 ...]
# Graphs and the stream fork are switched off so that every kernel is a visible, serialised launch (shares, not absolutes).
mkdir -p gpurun_out
for W in "$@"; do
  # launch list: steps 200..240 (11 launches per step without a rebuild)
  B200MD_USE_GRAPH=0 B200MD_NO_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 440 --csv \
      --log-file gpurun_out/r02_launches_$W.csv python tools/gpu_steps.py $W 60 200 > gpurun_out/r02_launches_$W.log 2>&1
  # one full capture of every kernel of ~5 consecutive steady-state steps (at least one with a list rebuild)
  B200MD_USE_GRAPH=0 B200MD_NO_OVERLAP=1 ncu --set full --clock-control none --import-source on -s 2600 -c 66 \
      -o gpurun_out/r02_kernels_$W python tools/gpu_steps.py $W 60 200 > gpurun_out/r02_kernels_$W.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
