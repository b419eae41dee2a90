"""Prints the headline metrics and the top stall sites of every kernel in an .ncu-rep."""
import csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sectors_op_red.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
names = []
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    names.append(name)
    print("##", name)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("  %-70s %s %s" % (w, r[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, hi in enumerate(his):
    h = rows[hi]
    end = his[k + 1] - 1 if k + 1 < len(his) else len(rows)
    body = [r for r in rows[hi + 1:end] if len(r) == len(h)]
    idx = {n: i for i, n in enumerate(h)}
    tot = sum(int(r[idx['# Samples']]) for r in body)
    stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    agg = {s: sum(int(r[idx[s]]) for r in body) for s in stalls}
    print("\n## kernel", k, rows[hi - 1][1] if hi > 0 else "", "samples", tot, "instrs", len(body))
    print("  ", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for r in sorted(body, key=lambda r: -int(r[idx['# Samples']]))[:topn]:
        st = sorted([(int(r[idx[s]]), s[6:]) for s in stalls], reverse=True)[:2]
        print("  %5s %7s  %-60s %s" % (r[idx['# Samples']], r[idx['Instructions Executed']], r[idx['Source']][:60], st))
