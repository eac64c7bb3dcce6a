"""profiles/<tag>_track_ncu_summary.csv + profiles/track_l1_traffic.json from an `ncu --page raw --csv` export of the four level kernels.
Usage: python tools/ncu_summary.py gpurun_out/<tag>_raw.csv <tag> "<command line the capture ran>" """
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
tag, cmd = sys.argv[2], sys.argv[3]
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed']
keep += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
ki = hdr.index('Kernel Name')
with open(f'profiles/{tag}_track_ncu_summary.csv', 'w') as f:
    f.write(f'# ncu --set full --clock-control none --import-source on -k regex:k_track_level, {cmd}; one column per level kernel\n')
    f.write('metric,unit,' + ','.join('"%s"' % r[ki].replace('void ', '').split('(')[0] for r in rows[2:]) + '\n')
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            f.write(f'{k},{units[i]},' + ','.join(r[i] for r in rows[2:]) + '\n')
i_r, i_w = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}  # ncu picks one unit per column
l1 = rows[-1]
n_problems = int(l1[hdr.index('launch__grid_size')]) // max(int(l1[hdr.index('launch__cluster_size')]), 1)
rd, wr = float(l1[i_r]) * UNIT[units[i_r]], float(l1[i_w]) * UNIT[units[i_w]]
json.dump({"kernel": l1[ki], "problems_per_launch": n_problems, "dram_bytes_per_problem": (rd + wr) / n_problems,
           "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
           "source": f"profiles/{tag}_track_ncu_summary.csv (ncu --set full, {cmd})"},
          open('profiles/track_l1_traffic.json', 'w'), indent=1)
