"""Per-SASS-instruction warp-stall samples of an .ncu-rep captured with --set full --import-source on: prints the stall
mix of the whole kernel and every instruction with at least `thr` samples (the mbarrier waits, wait::ld, MUFU queue ...).
usage: python tools/ncu_stall_hotspots.py gpurun_out/x.ncu-rep [thr]"""
import csv, sys, subprocess
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 900
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index("Source"); isamp = hdr.index("# Samples")
cols = {n: hdr.index(n) for n in ["stall_long_sb","stall_short_sb","stall_wait","stall_barrier","stall_math","stall_mio","stall_branch_resolving","stall_not_selected","stall_selected","stall_no_inst","stall_dispatch"]}
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot, "per warp", tot // 12)
agg = {}
for r in data:
    for k, v in cols.items():
        agg[k] = agg.get(k, 0) + int(r[v] or 0)
print({k: round(v / tot, 3) for k, v in agg.items()})
for i, r in enumerate(data):
    n = int(r[isamp] or 0)
    if n >= thr:
        st = {k: int(r[v] or 0) for k, v in cols.items()}
        st = {k: v for k, v in st.items() if v > 0.15 * n}
        print(i, n, r[isrc][:62], st)
