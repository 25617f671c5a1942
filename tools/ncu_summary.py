"""Summarise an .ncu-rep (read here, on the CPU box, with `ncu -i`) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.md ["title"]"""
import collections, csv, io, re, subprocess, sys

RAW = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'launch__grid_size', 'launch__block_size',
       'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
       'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
       'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
       'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg', 'sm__cycles_active.max', 'sm__cycles_elapsed.avg',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
       'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
       'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
       'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
       'smsp__sass_average_branch_targets_threads_uniform.pct']


def ncu(args):
    return subprocess.run(['ncu'] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on)", "",
             "## per-launch metrics", ""]
    names = [re.sub(r'\(.*', '', r[idx['Kernel Name']])[-40:] for r in data]
    lines.append("| metric | unit | " + " | ".join(f"{i}:{n}" for i, n in enumerate(names)) + " |")
    lines.append("|---|---|" + "---|" * len(names))
    for m in RAW:
        if m in idx:
            lines.append(f"| {m} | {units[idx[m]]} | " + " | ".join(r[idx[m]][:14] for r in data) + " |")
    srows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv']))))
    secs, cur = [], None
    for r in srows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': [], 'hdr': None}; secs.append(cur); continue
        if cur is None or not r:
            continue
        if r[0] == 'Address':
            cur['hdr'] = r; continue
        if cur['hdr']:
            cur['rows'].append(r)
    for k, s in enumerate(secs):
        h = {c: i for i, c in enumerate(s['hdr'])}
        tot = sum(int(r[h['Instructions Executed']] or 0) for r in s['rows'])
        ops = collections.Counter()
        for r in s['rows']:
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[h['Source']])
            ops[m.group(2).split('.')[0] if m else '?'] += int(r[h['Instructions Executed']] or 0)
        lines += ["", f"## launch {k}: SASS opcode mix ({len(s['rows'])} SASS instructions, {tot} warp-instructions executed)", ""]
        lines.append(", ".join(f"{op} {100 * c / max(tot, 1):.1f}%" for op, c in ops.most_common(16)))
        stalls = {c: sum(int(r[h[c]] or 0) for r in s['rows']) for c in s['hdr'] if c.startswith('stall_') and 'Not Issued' not in c}
        ts = sum(stalls.values()) or 1
        lines += ["", "warp stall samples: " + ", ".join(f"{c[6:]} {100 * v / ts:.1f}%" for c, v in sorted(stalls.items(), key=lambda x: -x[1])[:8])]
    open(out, 'w').write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == '__main__':
    main()
