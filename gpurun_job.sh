set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmm_stream_gpu.py tests/test_sparse_gpu.py tests/test_peer_gpu.py -x -q -m gpu > gpurun_out/test_new.log 2>&1; echo "new tests rc=$?"
tail -15 gpurun_out/test_new.log
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_edge_cases_gpu.py -x -q -m gpu > gpurun_out/test_models.log 2>&1; echo "model tests rc=$?"
tail -8 gpurun_out/test_models.log
timeout 400 python tools/shard_probe.py 8 > gpurun_out/shard_probe8.log 2>&1; echo "probe rc=$?"
tail -2 gpurun_out/shard_probe8.log | cut -c1-600
timeout 300 python tools/shard_ncu.py 8 > gpurun_out/shard_ncu_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'spmm_(chunk|stream)_kernel' --launch-skip 9 --launch-count 9 -o gpurun_out/shard_spmm_r02 -f python tools/shard_ncu.py 8 > gpurun_out/shard_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/shard_ncu.log
