# fp32 row-tile kernel with the balance-aware Q (45 maps: Q = 8, 15 warps): tests, throughput, one ncu --set full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32 or resident or empty_batch or weights_follow" > gpurun_out/r3d_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3d_pytest.log
B="python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs"
for model in res15 res15_narrow res26 res8; do
timeout 600 $B --model $model --batch 2048 > gpurun_out/r3d_bench_${model}.log 2> gpurun_out/r3d_bench_${model}.err
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3_f32_row -s 20 -c 1 -o gpurun_out/r3d_f32_row -f python bench.py --precision fp32 --model res15 --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs > gpurun_out/r3d_ncu.log 2>&1
echo finished
