"""Summarise ncu outputs into profiles/: (1) per-kernel time shares from a launch-list CSV, (2) key counters of one
--set full capture.  usage: python tools/ncu_summary.py launches.csv [prof.ncu-rep]"""
import collections, csv, re, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total ms | avg ms | share |", "|---|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.3f | %.1f%% |" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    return "\n".join(out)

WANT = [r"gpu__time_duration.sum$", r"sm__cycles_elapsed.avg.per_second", r"launch__registers_per_thread", r"launch__occupancy_limit_registers",
        r"sm__warps_active.avg.pct_of_peak_sustained_active", r"smsp__issue_active.avg.pct_of_peak_sustained_active",
        r"sm__inst_executed_pipe_(alu|fma|uniform|lsu)\.avg\.pct_of_peak_sustained_active", r"sm__pipe_(alu|fma)_cycles_active.avg.pct_of_peak_sustained_active",
        r"smsp__inst_executed.sum$", r"dram__bytes_(read|write).sum$", r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        r"smsp__average_warps_issue_stalled_(math_pipe_throttle|no_instruction|wait|dispatch_stall|not_selected|long_scoreboard|short_scoreboard)_per_issue_active",
        r"smsp__warps_eligible.avg.per_cycle_active", r"l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", r"lts__t_sector_hit_rate.pct"]

def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for val in rows[2:]:
        out.append("kernel: `%s`  grid %s block %s" % (val[hdr.index("Kernel Name")], val[hdr.index("Grid Size")], val[hdr.index("Block Size")]))
        out += ["| metric | value | unit |", "|---|---:|---|"]
        for i, h in enumerate(hdr):
            if any(re.search(w, h) for w in WANT):
                out.append("| %s | %s | %s |" % (h, val[i], units[i]))
    return "\n".join(out)

if __name__ == "__main__":
    print(launches(sys.argv[1]))
    if len(sys.argv) > 2:
        print()
        print(full(sys.argv[2]))
