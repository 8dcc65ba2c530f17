mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mfcc or stream or wave" > gpurun_out/r2l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest.log
timeout 120 python tools/fe_only.py > gpurun_out/r2l_fe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc_kernel -s 3 -c 1 -o gpurun_out/r2l_mfcc python tools/fe_only.py > gpurun_out/r2l_ncu.log 2>&1
echo finished
