mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_info.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "bf16x3" > gpurun_out/r2a_pytest_x3.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest_x3.log
timeout 1200 python -m pytest tests -m gpu -q -k "not bf16x3" > gpurun_out/r2a_pytest_rest.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest_rest.log
timeout 300 python tools/determinism_probe.py bf16 bf16x3 > gpurun_out/r2a_determinism.log 2>&1
HONK2_TC_SWEEP_DISCARD=0 timeout 300 python tools/determinism_probe.py bf16 > gpurun_out/r2a_determinism_nodiscard.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2>gpurun_out/r2a_bench.err
timeout 120 python tools/fe_only.py > gpurun_out/r2a_fe.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc_kernel -s 3 -c 1 -o gpurun_out/r2a_mfcc python tools/fe_only.py > gpurun_out/r2a_ncu.log 2>&1
echo finished
