"""profiles/: one-line-per-kernel table from .ncu-rep files (ncu --set full) + the per-kernel shares of a bench step
from the launch list (ncu --metrics gpu__time_duration.sum).  usage: python tools/ncu_table.py gpurun_out profiles/r1"""
import csv, io, json, re, subprocess, sys
from collections import defaultdict
from pathlib import Path
src, outp = Path(sys.argv[1]), sys.argv[2]
PEAK = json.loads(Path("MEASURED_PEAKS.json").read_text()) if Path("MEASURED_PEAKS.json").exists() else {}
def raw(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
def f(d, k, default=float("nan")):
    try: return float(d[k].replace(",", ""))
    except Exception: return default
lines = ["| capture | kernel | grid x block | time (us) | DRAM rd+wr (MB) | DRAM GB/s | of measured HBM peak | DRAM pipe % | tensor pipe % | XU (MUFU) % | issue active % | regs |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for rep in sorted(src.glob("r1_*.ncu-rep")):
    d, u = raw(rep)
    t_unit = u.get("gpu__time_duration.sum", "us")
    t = f(d, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(t_unit, 1)
    def mb(k):
        v = f(d, k); un = u.get(k, "byte")
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(un, 1e-6)
    mbs = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    gbs = mbs / 1e3 / (t * 1e-6) if t == t and t > 0 else float("nan")
    lines.append(f"| {rep.stem} | `{d.get('Kernel Name','?')[:44]}` | {d.get('Grid Size','?')} x {d.get('Block Size','?')} | {t:.1f} | {mbs:.1f} | {gbs:.0f} | "
                 f"{gbs / PEAK.get('hbm_gbs', 6573.5):.2f} | {f(d,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{f(d,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):.1f} | {f(d,'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed',0):.1f} | "
                 f"{f(d,'smsp__issue_active.avg.pct_of_peak_sustained_active',0):.1f} | {d.get('launch__registers_per_thread','?')} |")
Path(outp + "_ncu_kernels.md").write_text("# ncu --set full, one launch per kernel family (B200, --clock-control none)\n\n"
    "Times under ncu are cold-cache / serialised; use them for pipe utilisation and DRAM traffic, not as bench values.\n\n" + "\n".join(lines) + "\n")
# launch list -> shares
ll = src / "r1_bench_launches.csv"
if ll.exists():
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ui = hdr.index("Metric Unit")
    agg, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum": continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1}.get(r[ui], 1e-6)
        name = re.sub(r"<.*", "", r[ki].split("(")[0])
        agg[name] += v; cnt[name] += 1
    tot = sum(agg.values())
    out = ["# kernel shares of bench.py steps under `ncu --metrics gpu__time_duration.sum` (700 launches ~ 2 denoise steps)\n",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        out.append(f"| `{k[:60]}` | {cnt[k]} | {v:.2f} | {v / tot:.3f} |")
    Path(outp + "_bench_launch_shares.md").write_text("\n".join(out) + "\n")
print(Path(outp + "_ncu_kernels.md").read_text())
print(Path(outp + "_bench_launch_shares.md").read_text() if ll.exists() else "")
