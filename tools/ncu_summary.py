"""Summarise ncu outputs into profiles/: launch list -> compact csv + shares; --set full report -> transposed csv + key table."""
import collections, csv, json, subprocess, sys

def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]; ik, iv, ig, ib = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
    agg = collections.OrderedDict()
    with open(dst, 'w') as f:
        f.write("# %s\n# (cold-cache, serialised: compare SHARES with bench.py's live CUDA-event numbers, not absolutes)\n" % cmd)
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in rows[1:]:
            name = r[ik].split('(')[0].replace('void ', '')[:60]
            f.write("%s,%s,%s,%s,%s\n" % (r[0], name, r[ig].replace(',', ' '), r[ib].replace(',', ' '), r[iv]))
            agg.setdefault(name + " grid" + r[ig].replace(',', 'x').replace(' ', ''), []).append(float(r[iv].replace(',', '')))
        ours = {k: v for k, v in agg.items() if k.startswith('xsup::')}
        tot = sum(sum(v) for v in ours.values())
        f.write("# --- xsup kernels only, per (kernel, grid): the parity gate's small launches come first and are listed apart; name, launches, mean_us, share\n")
        for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
            f.write("# %s,%d,%.1f,%.4f\n" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
            print("%-50s n=%3d mean=%9.1f us share=%5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg', 'sm__cycles_active.min', 'sm__cycles_active.max', 'sm__cycles_elapsed.avg',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']

def full(rep, dst, cmd):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); h = rows[0]
    names = [r[h.index('Kernel Name')].split('(')[0] for r in rows[2:]]
    with open(dst, 'w') as f:
        f.write("# %s\nmetric,unit,%s\n" % (cmd, ",".join('"%s"' % n for n in names)))
        for i, name in enumerate(h):
            f.write('"%s","%s",%s\n' % (name, rows[1][i], ",".join('"%s"' % r[i] for r in rows[2:])))
    out = {}
    for r, n in zip(rows[2:], names):
        print("-----", n)
        for k in KEYS:
            if k in h:
                print("  %-80s %s %s" % (k, r[h.index(k)], rows[1][h.index(k)]))
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
        out[n] = int(sum(float(r[h.index(m)].replace(',', '')) * scale[rows[1][h.index(m)]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum')))
    return out

if __name__ == '__main__':
    what = sys.argv[1]
    if what == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        print(json.dumps(full(sys.argv[2], sys.argv[3], sys.argv[4])))
