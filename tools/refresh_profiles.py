"""Turn the raw captures in gpurun_out/ (see profiles/README.md for the commands) into the committed summaries."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "launches_r01_s2.csv"), os.path.join(P, "launches_r01_s2_bench.csv"))
rep = os.path.join(G, "prof_round_final.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(G, "round_final_raw.csv"), "w").write(raw)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
open(os.path.join(G, "round_final_src.csv"), "w").write(src)
lst = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_profile.py"), os.path.join(G, "round_final_src.csv"), "0", "32768"],
                     capture_output=True, text=True).stdout.split("\n")
keep = [l[:110] for i, l in enumerate(lst) if i < 2 or (len(l.split()) > 2 and float(l.split()[1]) > 0.5)]
open(os.path.join(P, "sass_q_round_kernel_r01_s2.txt"), "w").write("\n".join(keep) + "\n")
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'lts__t_sectors.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__warps_eligible.avg.per_cycle_active']
want += [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
with open(os.path.join(P, "ncu_round_r01_s2.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + ['launch %d' % i for i in range(len(rows) - 2)])
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
i, j = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
tot = [float(r[i]) * mul[units[i]] + float(r[j]) * mul[units[j]] for r in rows[2:]]
d = json.load(open(os.path.join(G, "bench_final.json")))
open(os.path.join(P, "bench_r01_s2_final.json"), "w").write(json.dumps(d) + "\n")
json.dump({"round_dram_bytes": sum(tot),
           "per_kernel": {"q_round_kernel<2,0>": tot[0], "q_pull_dense_kernel (sub-step 0)": tot[1], "q_pull_dense_kernel (sub-step 1)": tot[2]},
           "algorithmic_bytes_per_round": 52.0 * d["roofline"]["units_per_launch"],
           "source": "profiles/ncu_round_r01_s2.csv: dram__bytes_read.sum + dram__bytes_write.sum of the three kernels of one round, "
                     "ncu --set full --clock-control none on `python bench.py --steps 30 --warmup 5 --cpu-seconds 0 --no-aux` (cold L2 per "
                     "replay), 2^20 envs, layout 3 block 2, 2 cars. Records (16 MiB) are read from DRAM once per round and written back to "
                     "L2 only; the Q table, the state-info words and the per-CTA bitmaps are L2-resident."},
          open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
t = [r[hdr.index('gpu__time_duration.sum')] for r in rows[2:]]
n = [r[hdr.index('smsp__inst_executed.sum')] for r in rows[2:]]
print("kernel us", t, "warp-instructions", n, "dram bytes", tot)
print("bench value %.3e ms/step %.5f e2e %.3e frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
print({k: (v.get("value"), v.get("us_per_round")) for k, v in d["aux"].items()})
print(d["roofline"]["dominant_kernel"], d["gpu_launches"], d["clocks"])
