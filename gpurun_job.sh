python tools/tf32_peak.py > gpurun_out/tf32_peak.json 2> gpurun_out/tf32_peak.err; cat gpurun_out/tf32_peak.json; tail -2 gpurun_out/tf32_peak.err
