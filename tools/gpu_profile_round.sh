#!/bin/bash
# One GPU-box visit for the profile evidence of a round (run under gpurun from the repo root): $1 = tag.
#   1. the bench command without ncu (must exit 0), then its ncu launch list;
#   2. ncu --set full of one 512-board launch of every 16-bit kernel (after the same program ran without ncu), summaries + per-line tables;
#   3. the same for one 256-board exact-mode (fp32_split) forward.
# ONLY=split skips 1 and 2.  Only text leaves the box (the .ncu-rep files are tens of MB).
T=${1:-rXX}
mkdir -p gpurun_out
if [ "${ONLY:-all}" != "split" ]; then
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_plain.log 2>&1 || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_ncu.log 2>&1
python tools/gpu_profile_run.py 512 2 > gpurun_out/${T}_profile_plain.log 2>&1 || { echo "profile run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k 'regex:frontend3|stage[BCD]_kernel|global_head' --launch-skip 24 --launch-count 12 \
    -o /tmp/${T}_fused -f python tools/gpu_profile_run.py 512 2 > gpurun_out/${T}_ncu_fused.log 2>&1
python tools/ncu_summary.py /tmp/${T}_fused.ncu-rep > gpurun_out/${T}_fused_kernels_ncu_summary.txt 2>&1
python tools/ncu_lines.py /tmp/${T}_fused.ncu-rep frontend3 40 > gpurun_out/${T}_frontend3_source_lines.txt 2>&1
python tools/ncu_lines.py /tmp/${T}_fused.ncu-rep stageC 40 > gpurun_out/${T}_stageC_source_lines.txt 2>&1
python tools/ncu_lines.py /tmp/${T}_fused.ncu-rep stageD 40 > gpurun_out/${T}_stageD_source_lines.txt 2>&1
python tools/ncu_lines.py /tmp/${T}_fused.ncu-rep stageB 30 > gpurun_out/${T}_stageB_source_lines.txt 2>&1
fi
python tools/gpu_profile_split.py 256 > gpurun_out/${T}_split_plain.log 2>&1 || { echo "split run failed"; exit 1; }
ncu --profile-from-start off --set full --clock-control none --import-source on -o /tmp/${T}_split -f python tools/gpu_profile_split.py 256 > gpurun_out/${T}_ncu_split.log 2>&1
python tools/ncu_summary.py /tmp/${T}_split.ncu-rep > gpurun_out/${T}_split_ncu_summary.txt 2>&1
python tools/ncu_lines.py /tmp/${T}_split.ncu-rep dense_b00 30 > gpurun_out/${T}_split_b00_source_lines.txt 2>&1
python tools/ncu_lines.py /tmp/${T}_split.ncu-rep depthwise_x2_smem 30 > gpurun_out/${T}_split_depthwise_source_lines.txt 2>&1
python tools/ncu_lines.py /tmp/${T}_split.ncu-rep pointwise_x2 30 > gpurun_out/${T}_split_pointwise_source_lines.txt 2>&1
WAVES=1024 python tools/gpu_split_layers.py > gpurun_out/${T}_split_layers.txt 2>&1
ls -la /tmp/${T}_*.ncu-rep; wc -l gpurun_out/${T}_*
