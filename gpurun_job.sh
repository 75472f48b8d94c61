set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_peer_gpu.py -x -q -m gpu > gpurun_out/test_peer.log 2>&1; echo "peer tests rc=$?"
tail -15 gpurun_out/test_peer.log
timeout 600 python tools/shard_probe.py 8 > gpurun_out/shard_probe8.log 2>&1; echo "probe rc=$?"
tail -3 gpurun_out/shard_probe8.log
