"""Condenses ncu outputs into the small tracked files under profiles/.
    python tools/ncu_summary.py launches <launches.csv>            -> per-kernel time shares
    python tools/ncu_summary.py full <report.ncu-rep> [n_top] [kernel_regex]  -> key metrics + hottest SASS lines
                                                   (of the first captured launch whose name matches kernel_regex)
"""
import collections, csv, io, re, subprocess, sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {ms:.2f} | {100 * ms / tot:.1f}% |")
    print(f"| total | | {tot:.2f} | |")


def full(rep, n_top=12, kernel=None):
    sel = ["--kernel-name", "regex:" + kernel, "--launch-count", "1"] if kernel else []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    print("| metric | value | unit |\n|---|---:|---|")
    for k in keys:
        if k in d:
            print(f"| {k} | {d[k][1]} | {d[k][0]} |")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[ci["# Samples"]] or 0) for r in data)
    print(f"\nHottest SASS lines (of {tot:.0f} warp samples):\n\n| samples | share | executed | SASS | top stall |\n|---:|---:|---:|---|---|")
    for r in sorted(data, key=lambda r: -float(r[ci["# Samples"]] or 0))[:n_top]:
        s = float(r[ci["# Samples"]])
        top = max(((float(r[ci[h]] or 0), h[6:]) for h in stall))
        print(f"| {s:.0f} | {100 * s / tot:.1f}% | {float(r[ci['Instructions Executed']]):.0f} | `{r[ci['Source']].strip()[:70]}` | {top[1]} |")


def lines(rep, n_top=25, kernel=None):
    """Warp-stall samples per CUDA source line (needs -lineinfo): where a kernel spends its time."""
    sel = ["--kernel-name", "regex:" + kernel, "--launch-count", "1"] if kernel else []
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel,
                         capture_output=True, text=True).stdout
    fname, out = None, []
    for row in csv.reader(io.StringIO(src)):
        if len(row) == 2 and row[0] == "File Path":
            fname = row[1].split("/")[-1]
        elif len(row) > 8 and row[2] == "-" and row[0].isdigit():      # a source-line aggregate row
            try:
                out.append((float(row[6] or 0), float(row[7] or 0), fname, int(row[0]), row[1].strip()))
            except ValueError:
                pass
    tot = sum(o[0] for o in out) or 1.0
    print(f"| samples | share | warp instr | line | source |\n|---:|---:|---:|---|---|")
    for smp, ins, f, ln, text in sorted(out, key=lambda o: -o[0])[:n_top]:
        print(f"| {smp:.0f} | {100 * smp / tot:.1f}% | {ins:.0f} | {f}:{ln} | `{text[:90]}` |")
    print(f"| {tot:.0f} | 100% | {sum(o[1] for o in out):.0f} | total | |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "lines":
        lines(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25, sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 12, sys.argv[4] if len(sys.argv) > 4 else None)
