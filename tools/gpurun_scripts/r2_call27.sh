# ncu --set full records (final sweep kernel, narrow fp32 row kernel) + smoke; compute-sanitizer is closed on this pool (the call answered rc 86)
mkdir -p gpurun_out




timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3g_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r3g_smoke.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:resnet_tc_sweep -s 3 -c 1 -o gpurun_out/r3g_sweep -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs > gpurun_out/r3g_ncu_sweep.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3_f32_row -s 20 -c 1 -o gpurun_out/r3g_f32_narrow -f python bench.py --precision fp32 --model res15_narrow --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs > gpurun_out/r3g_ncu_narrow.log 2>&1
echo finished
