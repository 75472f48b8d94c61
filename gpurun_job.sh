mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"attention_wseg_kernel" -s 2 -c 1 -f -o gpurun_out/ncu_wseg_v3 python bench.py --workload attention --steps 1 --warmup 1 --eager > gpurun_out/ncu_w.log 2>&1; tail -1 gpurun_out/ncu_w.log
