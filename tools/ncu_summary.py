"""Condense an .ncu-rep into the text summary kept under profiles/ (details page, trimmed, plus the raw counters the
roofline uses: duration, DRAM bytes, pipe utilisation, stall samples).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<command that was profiled>" [--bytes cfgN] > profiles/rNN_....txt

--bytes cfgN also records dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel in the report under
profiles/ncu_dram_bytes.json[cfgN][<kernel base name>] -- the `roofline.traffic` bench.py reports for that config.
"""
import csv
import io
import subprocess
import sys

rep, cmdline = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
bytes_cfg = sys.argv[sys.argv.index("--bytes") + 1] if "--bytes" in sys.argv else None
KEEP = ("GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis", "Scheduler Statistics",
        "Warp State Statistics", "Instruction Statistics", "Launch Statistics", "Occupancy")
print(cmdline)
txt = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
kernel, keep = None, False
for line in txt.splitlines():
    s = line.strip()
    if line.startswith("  ") and not line.startswith("    ") and "(" in s and "Context" in s:
        kernel = s.split(" (")[0]
        continue
    if s.startswith("Section: "):
        name = s[len("Section: "):]
        keep = name in KEEP
        if keep:
            print("\n[%s] %s" % (kernel, name))
        continue
    if keep and s and not s.startswith("-") and not s.startswith("Metric Name") and not s.startswith("OPT") and not s.startswith("INF"):
        if line.startswith("    ") and len(s) < 140:
            print("    " + s)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
print("\n== raw counters per launch ==")
for r in rows[2:]:
    print("[%s]" % r[hdr.index("Kernel Name")])
    for k in WANT:
        if k in hdr:
            print("    %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), int(r[i])) for i, h in enumerate(hdr)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and r[i].isdigit()]
    tot = sum(v for _, v in st) or 1
    print("    warp-state samples: " + ", ".join("%s %.1f%%" % (n, 100.0 * v / tot) for n, v in sorted(st, key=lambda x: -x[1]) if v * 200 > tot))

if bytes_cfg:
    import json, os, re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_dram_bytes.json")
    try:
        db = json.load(open(path))
    except Exception:
        db = {}
    per = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        base = re.sub(r"^void\s+", "", name).split("<")[0].split("(")[0].split("::")[-1]

        def val(k):
            v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)].lower()
            return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        per.setdefault(base, []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    db.setdefault(bytes_cfg, {})
    for base, v in per.items():
        db[bytes_cfg][base] = {"bytes": sum(v) / len(v), "launches": len(v), "source": os.path.basename(rep) + ": " + cmdline}
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
