cd /root/repo
for bw in 24 32; do echo "== 1024-thread build, block_warps=$bw"; LRA_B200_LIB=$PWD/lra_b200/liblra_b200_t1024.so LRA_B200_MAP_BLOCK_WARPS=$bw LRA_B200_MAP_ARENA_MB=24 python tools/map_timing.py --preset ont --reads 8192 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads"; done
echo "== 16384 reads, 512-thread build bw=16"; LRA_B200_MAP_ARENA_MB=24 python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads"
echo "== 16384 reads, 1024-thread build bw=32"; LRA_B200_LIB=$PWD/lra_b200/liblra_b200_t1024.so LRA_B200_MAP_BLOCK_WARPS=32 LRA_B200_MAP_ARENA_MB=24 python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|map_reads"
