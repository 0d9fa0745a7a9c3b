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
            "(gradient accumulation adds, layout copies, tiny reductions of the losses, cuBLAS GEMMs of the two Linear layers and of the upsample-conv weight transforms).", ""]
    return "\n".join(out)


def raw_rows(path):
    """rows of the ncu raw page: from the report itself, or from the `<name>.raw.csv` exported on the GPU box
    (gpurun brings back at most 64 MiB, the reports with imported source are larger)."""
    if os.path.exists(path):
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        with open(path.replace(".ncu-rep", ".raw.csv")) as f:
            raw = f.read()
    return list(csv.reader(io.StringIO(raw)))


def have(path):
    return os.path.exists(path) or os.path.exists(path.replace(".ncu-rep", ".raw.csv"))


def full_table(path, title):
    rd = raw_rows(path)
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
         "Workload: MCGAN CIFAR10-shaped, batch 512, bf16 (`tools/profile_step.py`), one B200.  Commands: "
         + ("`tools/ncu_pass.sh`." if tag == "r01" else "`tools/ncu_pass_r02.sh` (the reports are exported to CSV on the GPU box: `gpurun_out/prof_*.raw.csv`)."),
         ""]
for name, title in (("launches_d.csv", "Launch list — one D update (D(real), G(z), D(G.detach()), backward, Adam)"),
                    ("launches_g.csv", "Launch list — one G update (G(z), D(G), backward, Adam)")):
    p = os.path.join(SRC, name)
    if os.path.exists(p):
        parts.append(launch_table(p, title))
for name, title in (("prof_conv_fwd.ncu-rep", "`--set full`: conv_fwd_kernel<256, 1> (generator, 256-wide) launches"),
                    ("prof_conv_fwd_big.ncu-rep", "`--set full`: conv_fwd_kernel<256, 1>, the 32x32 / 16x16 launches of a G update"),
                    ("prof_conv_fwd_d.ncu-rep", "`--set full`: conv_fwd_kernel<128, 2> (discriminator, 128-wide) launches"),
                    ("prof_wgrad.ncu-rep", "`--set full`: conv_wgrad_kernel, generator launches"),
                    ("prof_wgrad_d.ncu-rep", "`--set full`: conv_wgrad_kernel, discriminator launches"),
                    ("prof_gate.ncu-rep", "`--set full`: gate_fwd_kernel / gate_bwd_reduce_kernel / gate_bwd_apply_kernel, G update"),
                    ("prof_gate_fwd.ncu-rep", "`--set full`: bn_partial_kernel / gate_fwd_kernel, generator launches"),
                    ("prof_gate_bwd.ncu-rep", "`--set full`: gate_bwd_reduce_kernel / gate_bwd_apply_kernel"),
                    ("prof_thin.ncu-rep", "`--set full`: thin (RGB-side) conv kernels of a D update"),
                    ("prof_thin_wg.ncu-rep", "`--set full`: thin_conv_wgrad kernels")):
    p = os.path.join(SRC, name)
    if have(p):
        parts.append(full_table(p, title))

# per-launch DRAM traffic of the dominant shapes (bench.py attaches it to its roofline objects):
# (report, index of the launch inside the report) -> shape descriptor used by ops.PROFILE / bench.py
TRAFFIC = {"conv 512x32x32 256->256 k3 f0x5 bf16": ("prof_conv_fwd.ncu-rep", 1),
           "conv_ex kind2 512x16x16 256->256 f0x1": ("prof_conv_fwd.ncu-rep", 0),
           "conv 512x32x32 128->128 k3 f0x3 bf16": ("prof_conv_fwd_d.ncu-rep", 0),
           "thin_conv 512x32x32 3->128 k3 gate1": ("prof_thin.ncu-rep", 0),
           "gate_fwd 512x32x32x256 f0xb raw0": ("prof_gate_fwd.ncu-rep", 3),
           "bn_stats P524288 C256": ("prof_gate_fwd.ncu-rep", 2),
           "wgrad 512x32x32 256->256 k3": ("prof_wgrad.ncu-rep", 0)}
import json
if tag != "r01":
    # round 2+: the launch of a shape is found by its duration window [lo, hi] us inside the report (the capture
    # order depends on the step), not by a fixed index
    TRAFFIC = {"conv+bnstats kind0 512x32x32 256->256 f0x5": ("prof_conv_fwd_big.ncu-rep", (330, 480)),
               "conv+bnstats kind2 512x16x16 256->256 f0x1": ("prof_conv_fwd_big.ncu-rep", (180, 260)),
               "conv 512x32x32 128->128 k3 f0x3 bf16": ("prof_conv_fwd_d.ncu-rep", (105, 175)),
               "thin_conv 512x32x32 3->128 k3 gate1": ("prof_thin.ncu-rep", (44, 60)),
               "gate_fwd 512x32x32x256 f0xb raw0": ("prof_gate.ncu-rep", (90, 110)),
               "wgrad 512x32x32 256->256 k3": ("prof_wgrad.ncu-rep", (330, 600)),
               "wgrad 512x32x32 128->128 k3": ("prof_wgrad_d.ncu-rep", (105, 185))}
traffic = {}
for desc, (rep, idx) in TRAFFIC.items():
    pth = os.path.join(SRC, rep)
    if not have(pth):
        continue
    rd = raw_rows(pth)
    if isinstance(idx, tuple):
        hdr, units = rd[0], rd[1]
        di = hdr.index("gpu__time_duration.sum")
        scale = {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3}.get(units[di].lower(), 1.0)
        cand = [k for k, r in enumerate(rd[2:]) if idx[0] <= float(r[di].replace(",", "")) * scale <= idx[1]]
        if not cand:
            continue
        idx = cand[0]
    if len(rd) < 3 + idx:
        continue
    hdr, units, row = rd[0], rd[1], rd[2 + idx]

    def val(m):
        i = hdr.index(m)
        v, u = float(row[i].replace(",", "")), units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9,
                    "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1.0)
    traffic[desc] = {"kernel": row[hdr.index("Kernel Name")][:60], "ncu_duration_us": val("gpu__time_duration.sum") * 1e6,
                     "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                     "tensor_pipe_active_pct": float(row[hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]),
                     "source": f"{rep} launch {idx}, ncu --set full --clock-control none"}
with open(os.path.join(OUT, f"{tag}_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print("wrote", os.path.join(OUT, f"{tag}_traffic.json"))
with open(os.path.join(OUT, f"{tag}_ncu_summary.md"), "w") as f:
    f.write("\n".join(parts))
print("wrote", os.path.join(OUT, f"{tag}_ncu_summary.md"))
