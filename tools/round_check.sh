#!/bin/bash
# One GPU-box call that settles a tuning round (gpurun -- bash tools/round_check.sh TAG):
#   1. same-box A/B of the product library with programmatic dependent launch off / adaptive /
#      forced, plus any prebuilt variants in wrach_b200/lib/sweep/ (tools/ab.py: frame time and a
#      checksum of indices and positions after the same frames),
#   2. the whole GPU parity suite, 3. bench lines, 4. the ncu launch list of the bench command.
# Everything lands in gpurun_out/TAG_*; each phase has its own time limit.
TAG=${1:-r1b}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
extra=""
for so in wrach_b200/lib/sweep/lib_*.so; do [ -f "$so" ] && extra="$extra $(basename $so .so | sed s/lib_//)=$so"; done
( time timeout 240 python tools/ab.py 16m 200 2 off=:WRACH_PDL=0 product= $extra ) > gpurun_out/${TAG}_ab_16m.txt 2>&1
cat gpurun_out/${TAG}_ab_16m.txt
( timeout 90 python tools/ab.py 1m 500 2 off=:WRACH_PDL=0 product= forced=:WRACH_PDL=2 ) > gpurun_out/${TAG}_ab_1m.txt 2>&1
cat gpurun_out/${TAG}_ab_1m.txt
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.txt 2>&1
tail -5 gpurun_out/${TAG}_pytest.txt
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_16m.json 2> gpurun_out/${TAG}_bench_16m.err
cut -c1-300 gpurun_out/${TAG}_bench_16m.json
timeout 120 python bench.py --workload 1m --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_1m.json 2>/dev/null
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches_16m.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
