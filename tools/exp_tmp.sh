cd /root/repo
ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:ir_dp_warp_kernel -c 1 -f -o gpurun_out/r02y_irw python tools/map_timing.py --preset ont --reads 4096 --reps 1 --no-ref > gpurun_out/r02y_irwarp.log 2>&1
ls -la gpurun_out/r02y_irw.ncu-rep
