cd /root/repo
for lr in 0 192 2000; do echo "== LRA_B200_IR_LONG_ROWS=$lr"; LRA_B200_IR_LONG_ROWS=$lr python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|ir_dp|ir_band"; done
echo "== serial classes"; LRA_B200_SERIAL=1 python tools/map_timing.py --preset ont --reads 16384 --reps 2 --no-ref 2>&1 | grep -E "rep 1|ir_dp|ir_band"
