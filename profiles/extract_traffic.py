"""profiles/traffic.json from `ncu --set full` reports: DRAM bytes (read + write) per launch of each workload's dominant kernel.
usage: python profiles/extract_traffic.py  (reads gpurun_out/*.ncu-rep named below; bench.py reads the JSON)"""
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORTS = {                      # bench workload key -> report (the command lines are in profiles/README.md)
    'attention': 'gpurun_out/prof_gemm_tc_v3.ncu-rep',
    'attention_pool': 'gpurun_out/prof_attseg_v3.ncu-rep',
    'graph': 'gpurun_out/prof_spmm_v4.ncu-rep',
    'allpairs': 'gpurun_out/prof_allpairs_v2.ncu-rep',
    'k2_hbm': 'gpurun_out/prof_k2hbm_v2.ncu-rep',        # attention_wseg_tma_kernel, bench.py --workload k2hbm
    'k3_hbm': 'gpurun_out/prof_k3hbm_v1.ncu-rep',        # spmm_chunk_kernel, bench.py --workload k3hbm
}
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def dram_bytes(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    tot = []
    for r in rows[2:]:
        b = 0.0
        for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(name)
            b += float(r[i].replace(',', '')) * UNIT[units[i]]
        tot.append(b)
    return sum(tot) / len(tot), rows[2][hdr.index('Kernel Name')]


def main():
    out = os.path.join(ROOT, 'profiles', 'traffic.json')
    res = json.load(open(out)) if os.path.exists(out) else {}       # reports of earlier sessions may be gone: keep their numbers
    for key, rel in REPORTS.items():
        p = os.path.join(ROOT, rel)
        if os.path.exists(p):
            b, k = dram_bytes(p)
            res[key] = int(b)
            print(f'{key}: {b / 1e6:.1f} MB per launch  ({k[:60]})')
    json.dump(res, open(out, 'w'), indent=1)


if __name__ == '__main__':
    main()
