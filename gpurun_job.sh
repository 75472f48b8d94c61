mkdir -p gpurun_out
echo "== allpairs tests"; timeout 240 python -m pytest tests/test_allpairs_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "== allpairs bench"; timeout 120 python tools/allpairs_bench.py 4736 100000 2>&1 | tail -4
