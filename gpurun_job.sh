set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/bench3.err
