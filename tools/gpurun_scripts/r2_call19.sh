mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "permutation or gain_shift" > gpurun_out/r2u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_pytest.log
echo finished
