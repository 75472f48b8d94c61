mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "rc=$?"; tail -3 gpurun_out/bench4.err
