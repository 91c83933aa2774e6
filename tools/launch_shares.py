"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py (cold-cache, serialised
launches: the SHARES are the evidence, not the absolute times).  Splits the list at the attention launches into the part
that belongs to the DiT forwards and the VAE part around them.
usage: python tools/launch_shares.py gpurun_out/r2s_bench_launches.csv profiles/r2_bench_launch_shares.md [bench.json]"""
import csv
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows:
    if r is hdr or len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)            # -> ms
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name).replace("aether::", "")
    launches.append((name, v))

DIT = ("attention_v3", "attention_combine", "gemm2_kernel", "gemm_kernel", "ln_modulate", "qk_norm_rope", "small_m_linear",
       "patchify", "unpatchify", "timestep_sinusoid", "cfg_dpm_step")
first = next((i for i, (n, _) in enumerate(launches) if "ln_modulate" in n or "patchify" in n or "timestep" in n), 0)
last = max((i for i, (n, _) in enumerate(launches) if "cfg_dpm_step" in n or "unpatchify" in n), default=len(launches) - 1)
dit = [(n, v) for n, v in launches[first:last + 1] if any(k in n for k in DIT)]
vae = launches[:first] + launches[last + 1:]


def table(ls):
    agg, cnt = defaultdict(float), defaultdict(int)
    for n, v in ls:
        agg[n] += v
        cnt[n] += 1
    tot = sum(agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for n in sorted(agg, key=lambda k: -agg[k]):
        out.append(f"| `{n[:70]}` | {cnt[n]} | {agg[n]:.2f} | {agg[n] / tot:.3f} |")
    return out, tot, agg


t_dit, tot_dit, agg_dit = table(dit)
t_vae, tot_vae, _ = table(vae)
n_fwd = sum(1 for n, _ in dit if "unpatchify" in n) or 1
attn = sum(v for n, v in agg_dit.items() if "attention" in n)
txt = [f"# kernel shares of `bench.py` under `ncu --metrics gpu__time_duration.sum` ({len(launches)} launches of this library's kernels: "
       f"the tail of the VAE encode, {n_fwd} DiT forward(s) + DPM steps, the head of the VAE decode; cold-cache, serialised: the SHARES "
       f"are the evidence, not the absolute times)", "", f"## DiT forwards ({len(dit)} launches, {tot_dit:.1f} ms under ncu = "
       f"{tot_dit / n_fwd:.1f} ms per forward)", ""] + t_dit
txt += ["", f"attention share of a forward under ncu: **{attn / tot_dit:.3f}**"]
if len(sys.argv) > 3 and Path(sys.argv[3]).exists():
    d = json.loads(Path(sys.argv[3]).read_text().strip().splitlines()[-1])
    rf = d["roofline"]
    fwd = d.get("gpu_library_baseline", {}).get("ours_dit_forward_ms")
    if fwd:
        txt.append(f"bench.py's CUDA-event figures for the same build (`{Path(sys.argv[3]).name}`): attention {rf['avg_launch_ms']:.3f} ms x 42 = "
                   f"{rf['avg_launch_ms'] * 42:.1f} ms of a {fwd:.1f} ms forward = {rf['avg_launch_ms'] * 42 / fwd:.3f} "
                   f"(share of the whole round incl. VAE: {rf['share_of_step']:.3f})")
txt += ["", f"## VAE kernels around them ({len(vae)} launches, {tot_vae:.1f} ms under ncu)", ""] + t_vae
Path(sys.argv[2]).write_text("\n".join(txt) + "\n")
print("\n".join(txt[:30]))
