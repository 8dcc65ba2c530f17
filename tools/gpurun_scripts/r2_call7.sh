mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:resnet_tc_sweep -s 2 -c 1 -o gpurun_out/r2g_narrow_sweep $B --model res15_narrow --precision bf16 > gpurun_out/r2g_ncu_narrow_sweep.log 2>&1
timeout 600 $NCU -k regex:conv3x3_f32 -s 20 -c 1 -o gpurun_out/r2g_narrow_fp32 $B --model res15_narrow --precision fp32 --batch 1024 > gpurun_out/r2g_ncu_narrow_fp32.log 2>&1
timeout 600 $NCU -k regex:cnn_tc_fused -s 2 -c 1 -o gpurun_out/r2g_cnn_fused $B --model cnn-trad-fpool3 --precision bf16 --chunk 8192 > gpurun_out/r2g_ncu_cnn.log 2>&1
timeout 600 $NCU -k regex:resnet_tc_sweep -s 2 -c 1 -o gpurun_out/r2g_res15_sweep $B > gpurun_out/r2g_ncu_res15.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_cnn.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --model cnn-trad-fpool3 --precision bf16 --chunk 8192 > gpurun_out/r2g_l2.log 2>&1
timeout 300 python tools/determinism_probe.py bf16 bf16x3 > gpurun_out/r2g_determinism.log 2>&1
echo finished
