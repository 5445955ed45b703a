#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU):
   ncu_summary.py launches <launches.csv>      per-kernel time shares of a launch list
   ncu_summary.py kernel <report.ncu-rep>      key metrics + opcode mix + stall reasons + hottest SASS"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ix = {k: j for j, k in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r[ix["Metric Value"]])
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}[r[ix["Metric Unit"]]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'total ms':>10s} {'n':>5s} {'share':>7s}  kernel   (cold-cache, serialised: compare shares)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.3f} {c:5d} {t / tot * 100:6.1f}%  {n}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def kernel(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:70s} {r[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= len(hdr) and r[0] != "Address":
            data.append(r)
    f = lambda r, k: float(r[ix[k]] or 0)  # noqa: E731
    tot_i = sum(f(r, "Instructions Executed") for r in data)
    tot_s = sum(f(r, "# Samples") for r in data)
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += f(r, "Instructions Executed")
        samp[op] += f(r, "# Samples")
    print(f"-- first kernel: {len(data)} SASS instructions, {tot_i:.3g} executed, {tot_s:.0f} samples")
    print("-- opcode mix (share of executed warp instructions / of stall samples)")
    for op, c in ops.most_common(14):
        print(f"  {op:10s} {c / tot_i * 100:6.2f}%  {samp[op] / tot_s * 100:6.2f}%")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter({s: sum(f(r, s) for r in data) for s in stalls})
    S = sum(tot.values())
    print("-- stall reasons")
    for s, c in tot.most_common(8):
        print(f"  {s:26s} {c / S * 100:6.2f}%")
    print("-- hottest SASS by samples")
    for i in sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:top]:
        r = data[i]
        st = sorted(((s, f(r, s)) for s in stalls), key=lambda kv: -kv[1])[0]
        print(f"  {f(r, '# Samples') / tot_s * 100:5.2f}% exec={f(r, 'Instructions Executed'):12.0f} "
              f"{r[ix['Source']][:64]:64s} {st[0]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
