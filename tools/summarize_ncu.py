#!/usr/bin/env python
"""Turn an `ncu --set full` report into the markdown summary + K1 DRAM-traffic JSON kept under profiles/.
usage: python tools/summarize_ncu.py gpurun_out/v5_full.ncu-rep profiles/r01_v5 "<command that was profiled>" """
import collections, csv, io, json, re, subprocess, sys

rep, prefix, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def tobytes(v, u):
    return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def short(n):
    return n.replace('void ', '').replace('jb::', '').split('(')[0]


lines = [f"# ncu --set full --clock-control none: {rep.split('/')[-1]}\n\n", f"Command: `{cmd}` (exited 0 without ncu first).\n",
         "ncu replays each kernel with caches flushed, so K2's coefficient reads show up as DRAM reads here; in a normal run they are L2 hits.\n"]
out = {}
for r in rows[2:]:
    name = short(r[hdr.index('Kernel Name')])
    if name in out:
        continue
    lines.append(f"\n## {name}\n\n| metric | value |\n|---|---|\n")
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f"| {w} | {r[i]} {units[i]} |\n")
            d[w] = (r[i], units[i])
    out[name] = d
if 'k_fused_blocks' in out:
    k1 = out['k_fused_blocks']
    rd, wr = tobytes(*k1['dram__bytes_read.sum']), tobytes(*k1['dram__bytes_write.sum'])
    json.dump({"kernel": "k_fused_blocks", "workload": "3840x2160 single image", "dram_bytes_per_launch": int(rd + wr),
               "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "source": f"{prefix}_ncu_full_summary.md"},
              open('profiles/k1_dram_traffic.json', 'w'), indent=1)
src = subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
seen = set()
for si in range(len(start) - 1):
    name = short(rows[start[si]][1])
    if name in seen:
        continue
    seen.add(name)
    seg = rows[start[si] + 1:start[si + 1]]
    h, data = seg[0], seg[1:]
    cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot, byop, ex = collections.Counter(), collections.Counter(), 0
    iE, iS = h.index('Instructions Executed'), h.index('Source')
    for r in data:
        e = int(r[iE] or 0)
        ex += e
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS].strip())
        byop[m.group(2) if m else '?'] += e
        for i, c in cols:
            try:
                tot[c] += int(r[i] or 0)
            except ValueError:
                pass
    s = sum(tot.values()) or 1
    lines.append(f"\n### {name}: warp instructions executed {ex:,}; stall samples: "
                 + ", ".join(f"{c[6:]} {100 * v / s:.0f}%" for c, v in tot.most_common(8)) + "\n\n")
    lines.append("Top opcodes (warp instructions): " + ", ".join(f"{o} {v:,}" for o, v in byop.most_common(14)) + "\n")
open(f"{prefix}_ncu_full_summary.md", 'w').writelines(lines)
print(''.join(lines))
