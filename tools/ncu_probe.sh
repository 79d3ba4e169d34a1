mkdir -p gpurun_out
B="python bench.py --steps 12 --warmup 10 --no-cpu-baseline --no-extra"
for sec in SpeedOfLight LaunchStats Occupancy SchedulerStats WarpStateStats InstructionStats MemoryWorkloadAnalysis ComputeWorkloadAnalysis SourceCounters; do
  timeout 120 ncu --section $sec --clock-control none -k regex:k_tile_frame -s 12 -c 1 -f -o gpurun_out/r2h_$sec $B > gpurun_out/r2h_$sec.log 2>&1
  echo "$sec: $(grep -c LaunchFailed gpurun_out/r2h_$sec.log) $(grep -E 'pass' gpurun_out/r2h_$sec.log | tail -1)"
done
