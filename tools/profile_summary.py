"""Summarise gpurun_out ncu artefacts into profiles/ (launch shares + key counters)."""
import collections
import csv
import json
import os
import subprocess
import sys

import json
import os

tag, kern = sys.argv[1], sys.argv[2]           # e.g. r2_int8 oz_score oz_kstar
kerns = sys.argv[2:]
# library kernel name + bench workload a capture stands for (profiles/traffic.json feeds bench.py's roofline.traffic)
TRAFFIC_KEYS = {"oz_score": ("rbf_n4096_d8_ei", "oz_score_kernel"), "score_gemm": ("rbf_n4096_d8_ei", "score_gemm_kernel")}
rows = [r for r in csv.reader(open('gpurun_out/launches.csv' if os.path.exists('gpurun_out/launches.csv') else 'profiles/%s_launches.csv' % tag)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split('(')[0]
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    v_us = v / 1000.0 if u.startswith('ns') else (v if u.startswith('us') else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v_us
tot = sum(a[1] for a in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none over a short bench.py run",
       "# (cold-cache, serialised launches: compare SHARES, not absolutes)", "kernel,launches,total_us,share"]
STEP = ('kstar', 'score_gemm', 'moments', 'acq_kernel', 'argmax_final', 'oz_score', 'oz_kstar', 'oz_moments', 'oz_flag', 'oz_sort',
        'oz_gather', 'oz_scatter', 'topk_pass', 'incumbent_')
step_tot = sum(t for k, (c, t) in agg.items() if any(s in k for s in STEP))
out[2] = "kernel,launches,total_us,share_of_all,share_of_scoring_step"
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    in_step = any(s in k for s in STEP)
    out.append("%s,%d,%.1f,%.4f,%s" % (k, c, t, t / tot, ("%.4f" % (t / step_tot)) if in_step else "-"))
open('profiles/%s_launch_shares.csv' % tag, 'w').write("\n".join(out) + "\n")
print("\n".join(out))
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.peak_sustained',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_src_fp64.avg.peak_sustained', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum']
for kern in kerns:
    if os.path.exists('gpurun_out/prof_%s_raw.csv' % kern):
        raw = open('gpurun_out/prof_%s_raw.csv' % kern).read()
    else:
        raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_%s.ncu-rep' % kern, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u, v = rr[0], rr[1], rr[2]
    lines = ["# ncu --set full --clock-control none -k regex:%s, first captured launch after the skip (short bench.py --quick run)" % kern, ""]
    for k in want:
        if k in h:
            i = h.index(k); lines.append("%s = %s %s" % (k, v[i], u[i]))
    open('profiles/%s_%s_ncu.txt' % (tag, kern), 'w').write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if kern in TRAFFIC_KEYS and 'dram__bytes_read.sum' in h:
        def to_bytes(name):
            i = h.index(name)
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[i]]
            return float(v[i].replace(',', '')) * mult
        path = 'profiles/traffic.json'
        t = json.load(open(path)) if os.path.exists(path) else {}
        wl, name = TRAFFIC_KEYS[kern]
        t.setdefault(wl, {})[name] = dict(dram_bytes=to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum'),
                                          dram_read=to_bytes('dram__bytes_read.sum'), dram_write=to_bytes('dram__bytes_write.sum'),
                                          source='profiles/%s_%s_ncu.txt (ncu --set full, one launch = 32768 candidates)' % (tag, kern))
        json.dump(t, open(path, 'w'), indent=1)
