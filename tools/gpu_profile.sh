#!/bin/bash
# Round-2 profile capture (run under gpurun, ONE GPU):  bash tools/gpu_profile.sh
# Graphs and the stream fork are switched off so that every kernel is a visible, serialised launch (shares, not absolutes).
# The .ncu-rep files are summarised ON the box (tools/summarize_profile_r02.py writes profiles/r02_*) and removed: gpurun_out/
# only carries 64 MiB back.
mkdir -p gpurun_out profiles
W=dhfr
B200MD_USE_GRAPH=0 B200MD_NO_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 220 --csv \
    --log-file gpurun_out/r02_launches_$W.csv python tools/gpu_steps.py $W 40 200 > gpurun_out/r02_launches_$W.log 2>&1
B200MD_USE_GRAPH=0 B200MD_NO_OVERLAP=1 ncu --set full --clock-control none -s 2600 -c 44 \
    -o gpurun_out/r02_kernels_$W python tools/gpu_steps.py $W 40 200 > gpurun_out/r02_kernels_$W.log 2>&1
python tools/summarize_profile_r02.py $W > gpurun_out/r02_summary_$W.log 2>&1
W=apoa1
B200MD_USE_GRAPH=0 B200MD_NO_OVERLAP=1 ncu --set full --clock-control none -k regex:'k_pair|k_pme_spread|k_pme_gather|k_fft' -s 1200 -c 6 \
    -o gpurun_out/r02_kernels_$W python tools/gpu_steps.py $W 30 200 > gpurun_out/r02_kernels_$W.log 2>&1
python tools/summarize_profile_r02.py $W > gpurun_out/r02_summary_$W.log 2>&1
rm -f gpurun_out/*.ncu-rep
cp profiles/r02_*_kernels.md profiles/r02_*_launch_summary.csv profiles/r02_*_k_pair_summary.json gpurun_out/ 2>/dev/null
cat gpurun_out/r02_summary_dhfr.log gpurun_out/r02_summary_apoa1.log
ls -la gpurun_out/
