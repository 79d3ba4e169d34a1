#!/bin/bash
# One GPU-box call that settles a tuning round (gpurun -- bash tools/round_check.sh TAG):
#   1. same-box A/B of the prebuilt variants in wrach_b200/lib/sweep/ and of the WRACH_PDL knob
#      (tools/ab.py: frame time + a checksum of indices and positions after the same frames),
#   2. the whole GPU parity suite with the candidate configuration,
#   3. bench lines of the candidate and the fallback, 4. the ncu launch list of the bench command.
# Everything lands in gpurun_out/TAG_*; each phase has its own time limit.
TAG=${1:-r1b}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=wrach_b200/lib/sweep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
( time timeout 240 python tools/ab.py 16m 200 2 base=$S/lib_base.so:WRACH_PDL=0 new=:WRACH_PDL=0 pdl=:WRACH_PDL=1 \
    noldcg=$S/lib_noldcg.so:WRACH_PDL=1 late=$S/lib_late.so:WRACH_PDL=1 ) > gpurun_out/${TAG}_ab_16m.txt 2>&1
cat gpurun_out/${TAG}_ab_16m.txt
( time WRACH_PDL=1 timeout 480 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_pdl1.txt 2>&1
tail -5 gpurun_out/${TAG}_pytest_pdl1.txt
WRACH_PDL=1 timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_16m_pdl1.json 2> gpurun_out/${TAG}_bench_16m_pdl1.err
WRACH_PDL=0 timeout 120 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_16m_pdl0.json 2> gpurun_out/${TAG}_bench_16m_pdl0.err
cut -c1-400 gpurun_out/${TAG}_bench_16m_pdl1.json
WRACH_PDL=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches_16m.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
( timeout 90 python tools/ab.py 1m 500 2 base=$S/lib_base.so:WRACH_PDL=0 new=:WRACH_PDL=0 pdl=:WRACH_PDL=1 ) > gpurun_out/${TAG}_ab_1m.txt 2>&1
cat gpurun_out/${TAG}_ab_1m.txt
WRACH_PDL=1 timeout 120 python bench.py --workload 1m --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_1m_pdl1.json 2>/dev/null
ls -la gpurun_out | tail -12
