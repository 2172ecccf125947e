# tools/run_driver_like.sh -- under gpurun: what the driver runs at round end, in its order
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python -m pytest tests/ -x -q -m gpu ) 2>&1 | tail -5
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
python bench.py --gpus 1 --steps 100 --warmup 3 2>/dev/null | tail -1 > gpurun_out/driver_like_bench.json; cut -c1-420 gpurun_out/driver_like_bench.json
ZC_SLOW=1 python -m pytest tests/test_gpu_testbench.py -x -q -m gpu 2>&1 | tail -2
