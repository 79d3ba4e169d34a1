#!/bin/bash
# GPU box, one GPU: the round's single-GPU records (gpurun -- bash tools/final_1gpu.sh TAG)
#   bench line with default flags, ncu launch list of the same command, one `ncu --set full` capture
#   of the dominant kernel.  Everything lands in gpurun_out/TAG_*.
TAG=${1:-r2f}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench_16m.json 2> gpurun_out/${TAG}_bench_16m.err
cut -c1-400 gpurun_out/${TAG}_bench_16m.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches_16m.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-extra > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_launches_16m.csv | cut -c1-220
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_frame -s 12 -c 1 -f \
    -o gpurun_out/${TAG}_full_16m python bench.py --steps 12 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out | tail -8
