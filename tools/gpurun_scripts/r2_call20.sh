mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity"
for mb in -1 0 48 80 110; do
HONK2_TC_L2_PERSIST_MB=$mb timeout 300 $B > gpurun_out/r2w_bench_$mb.log 2>gpurun_out/r2w_bench_$mb.err
done
for mb in -1 80; do
HONK2_TC_L2_PERSIST_MB=$mb timeout 600 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:resnet_tc_sweep -s 2 -c 1 --csv --log-file gpurun_out/r2w_ncu_$mb.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity > gpurun_out/r2w_ncu_$mb.log 2>&1
done
echo finished
