set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_gpu.py -x -q -m gpu > gpurun_out/test_peer.log 2>&1; echo "peer tests rc=$?"
tail -5 gpurun_out/test_peer.log
