set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t45.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/t45.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke45.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke45.log
( time python bench.py > gpurun_out/b45_full.json 2> gpurun_out/b45_full.err ) 2> gpurun_out/b45_time.txt; echo "bench rc=$?"; tail -3 gpurun_out/b45_full.err; cat gpurun_out/b45_time.txt
( time python bench.py --impl reference > gpurun_out/b45_ref.json 2> gpurun_out/b45_ref.err ) 2> gpurun_out/b45_ref_time.txt; echo "ref rc=$?"; cat gpurun_out/b45_ref_time.txt
