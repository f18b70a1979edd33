cd /root/repo
python -m pytest tests/test_sdp.py tests/test_map_e2e.py -q -m gpu -x 2>&1 | tail -3
LRA_B200_MAP_PROFILE=1 python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | tail -40 > gpurun_out/r02r_profile_5mb.log; tail -38 gpurun_out/r02r_profile_5mb.log | head -24
export LRA_B200_MAP_ARENA_MB=20
ncu --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:map_reads_kernel -c 1 --csv --log-file gpurun_out/r02r_map_dram.csv python tools/map_timing.py --preset ont --reads 8192 --reps 1 --no-ref > gpurun_out/r02r_map_dram.log 2>&1
grep -E "map_reads" gpurun_out/r02r_map_dram.csv | cut -c1-300
