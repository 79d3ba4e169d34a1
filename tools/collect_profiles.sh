#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/collect_profiles.sh TAG): bench line, ncu launch list of the
# same command, one `ncu --set full` capture of the step's kernels.  Results land in gpurun_out/.
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_16m.json 2> gpurun_out/${TAG}_bench_16m.err
tail -2 gpurun_out/${TAG}_bench_16m.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches_16m.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_phys|k_rebin|k_run_scan" -s 30 -c 3 \
    -o gpurun_out/${TAG}_full_16m python bench.py --steps 12 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -5
