mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_bench.log 2>gpurun_out/r2p_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2p_ref.log 2>gpurun_out/r2p_ref.err
echo finished
