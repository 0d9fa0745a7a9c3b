"""Turn the ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.
    python tools/summarize_profiles.py r01
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def launch_table(path, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        try:
            rows.append((r["Kernel Name"], float(r["Metric Value"])))
        except Exception:
            pass
    tot = sum(t for _, t in rows)
    if not rows:
        return f"### {title}\n\n(no launches captured)\n"
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, t in rows:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void ", "", k)[:90]
        agg[k][0] += 1
        agg[k][1] += t
    out = [f"### {title}", "", f"{len(rows)} launches, {tot / 1e6:.3f} ms of kernel time "
           "(ncu `gpu__time_duration.sum`, `--clock-control none`; cold-cache, serialised: compare SHARES).", "",
           "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append(f"| `{k}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.1f} % |")
    mine = sum(t for k, (n, t) in agg.items() if k.startswith("mcgm::"))
    out += ["", f"libmcgm kernels: {100 * mine / tot:.1f} % of kernel time; the rest are torch launches "
            "(Adam multi-tensor, tiny reductions, casts, cuBLAS GEMV/GEMM of the two Linear layers).", ""]
    return "\n".join(out)


def full_table(path, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    if len(rd) < 3:
        return f"### {title}\n\n(no data)\n"
    hdr, units = rd[0], rd[1]
    idx = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    out = [f"### {title}", "", "| kernel | " + " | ".join(f"{m.split('.')[0]} [{units[i]}]" for m, i in idx) + " |",
           "|---|" + "---:|" * len(idx)]
    ki = hdr.index("Kernel Name")
    for r in rd[2:]:
        out.append(f"| `{r[ki][:60]}` | " + " | ".join(r[i][:12] for _, i in idx) + " |")
    return "\n".join(out) + "\n"


os.makedirs(OUT, exist_ok=True)
parts = [f"# ncu summaries — round {tag}", "",
         "Workload: MCGAN CIFAR10-shaped, batch 512, bf16 (`tools/profile_step.py`), one B200.  Commands: `tools/ncu_pass.sh`.",
         ""]
for name, title in (("launches_d.csv", "Launch list — one D update (D(real), G(z), D(G.detach()), backward, Adam)"),
                    ("launches_g.csv", "Launch list — one G update (G(z), D(G), backward, Adam)")):
    p = os.path.join(SRC, name)
    if os.path.exists(p):
        parts.append(launch_table(p, title))
for name, title in (("prof_conv_fwd.ncu-rep", "`--set full`: conv_fwd_kernel, generator 256->256 3x3 @32x32 launches"),
                    ("prof_conv_fwd_d.ncu-rep", "`--set full`: conv_fwd_kernel, discriminator 128->128 launches"),
                    ("prof_wgrad.ncu-rep", "`--set full`: conv_wgrad_kernel, generator launches"),
                    ("prof_gate_fwd.ncu-rep", "`--set full`: bn_partial_kernel / gate_fwd_kernel, generator launches"),
                    ("prof_gate_bwd.ncu-rep", "`--set full`: gate_bwd_reduce_kernel / gate_bwd_apply_kernel")):
    p = os.path.join(SRC, name)
    if os.path.exists(p):
        parts.append(full_table(p, title))
with open(os.path.join(OUT, f"{tag}_ncu_summary.md"), "w") as f:
    f.write("\n".join(parts))
print("wrote", os.path.join(OUT, f"{tag}_ncu_summary.md"))
