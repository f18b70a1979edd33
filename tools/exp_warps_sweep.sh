cd /root/repo
out=gpurun_out/r02d_warps_sweep.log; : > $out
for w in 2 4 8 16; do
  echo "== warps_per_sm=$w" >> $out
  LRA_B200_MAP_WARPS_PER_SM=$w LRA_B200_MAP_ARENA_MB=24 python tools/map_timing.py --preset ont --reads 8192 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads" >> $out
done
cat $out
export LRA_B200_MAP_ARENA_MB=20
ncu --section SpeedOfLight --section SchedulerStats --section WarpStateStats --section Occupancy --clock-control none -k regex:map_reads_kernel -c 1 -f -o /tmp/mapfull python tools/map_timing.py --preset ont --reads 8192 --reps 1 --no-ref > gpurun_out/r02d_mapfull.log 2>&1
ncu -i /tmp/mapfull.ncu-rep --page raw --csv > gpurun_out/r02d_mapfull_raw.csv 2>/dev/null
tail -3 gpurun_out/r02d_mapfull.log
