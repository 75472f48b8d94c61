set -x
mkdir -p gpurun_out
timeout 300 python tools/shard_ncu.py 8 > gpurun_out/shard_ncu_plain.log 2>&1; echo "plain rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name regex:spmm_chunk_kernel --launch-skip 5 --launch-count 5 -o gpurun_out/shard_spmm_r02 -f python tools/shard_ncu.py 8 > gpurun_out/shard_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/shard_ncu.log
timeout 600 python -m pytest tests/test_sparse_gpu.py -x -q -m gpu > gpurun_out/test_sparse.log 2>&1; echo "sparse tests rc=$?"
tail -15 gpurun_out/test_sparse.log
