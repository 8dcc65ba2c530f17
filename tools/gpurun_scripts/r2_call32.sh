# final validation of the committed tree: whole GPU suite, smoke(), the driver's bench command
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r3l_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3l_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3l_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r3l_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r3l_bench_default.log 2> gpurun_out/r3l_bench_default.err
echo finished
