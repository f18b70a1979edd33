cd /root/repo
echo "== 768-thread build, block_warps=24"; LRA_B200_LIB=$PWD/lra_b200/liblra_b200_t768.so LRA_B200_MAP_BLOCK_WARPS=24 python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads"
echo "== 768-thread build, block_warps=20"; LRA_B200_LIB=$PWD/lra_b200/liblra_b200_t768.so LRA_B200_MAP_BLOCK_WARPS=20 python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads"
echo "== 512-thread build bw=16"; python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads"
echo "== 512-thread build bw=16, 32768 reads"; python tools/map_timing.py --preset ont --reads 32768 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads"
