python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime --no-train-step > gpurun_out/b51_att.json 2> gpurun_out/b51_att.err; echo "rc=$?"; tail -2 gpurun_out/b51_att.err
python bench.py --workload graph --no-cpu-baseline --no-train-step > gpurun_out/b51_graph.json 2> gpurun_out/b51_graph.err; echo "rc=$?"; tail -2 gpurun_out/b51_graph.err
