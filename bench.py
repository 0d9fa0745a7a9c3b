"""bench.py — MCGAN G+D training throughput (images/s) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host cores

A "step" is ONE reference training iteration (src/train_gan.py:139-176: 5 D updates + 1 G
update, hinge loss, Adam) of the CIFAR10-shaped MCGAN (32x32x3, 10 modes, controller rate
0.5, G widths 256x4, D widths 128x4, latent 128) at batch 512 per GPU, bf16 activations with
fp32 accumulation, synthetic data, random-init weights (seed 0).  One JSON line is printed by
rank 0.  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MCGAN G+D train imgs/s"
WORKLOAD = "MCGAN CIFAR10-shaped 32x32x3, 10 modes, control 0.5, one iteration = 5 D + 1 G updates (train_gan.py:139-176)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="images per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=64, help="images per step of the CPU reference arm (BASELINE configs[0])")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="time budget of the CPU reference arm's timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_reference(batch, steps, warmup, threads=None, budget_s=None):
    """The reference algorithm (oracle port: same torch CPU ops as the reference modules) on the
    host cores: up to `steps` iterations of 5D+1G at `batch` images (fewer if `budget_s` would be exceeded,
    judged from the first timed iteration).  Returns (img/s, cores, s/iter, iterations timed)."""
    import torch
    from mcgm_b200 import cfg, models, process_control
    from oracle import mcgan_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg.update(device="cpu", data_name="CIFAR10", model_name="mcgan", control={"controller_rate": "0.5"})
    process_control()
    cfg["classes_size"] = 10
    torch.manual_seed(0)
    sd = O.make_trainable(models.mcgan().state_dict())
    opt = O.make_adams(sd)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(batch, 3, 32, 32, generator=g) * 2 - 1
    label = torch.randint(0, 10, (batch,), generator=g)
    times = []
    i = 0
    while i < warmup + steps:
        zs = [torch.randn(batch, 128, generator=g) for _ in range(6)]
        t0 = time.perf_counter()
        O.train_iteration(sd, opt, img, label, zs, 10, [2, 1, 1])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            if budget_s is not None and sum(times) + dt > budget_s:
                break
        i += 1
    per = sum(times) / len(times)
    return batch / per, torch.get_num_threads(), per, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    value, cores, per, done = cpu_reference(args.ref_batch, steps, warmup, budget_s=args.ref_seconds)
    sample = (f"{done} iterations (5D+1G) at batch {args.ref_batch} (BASELINE configs[0]) after {warmup} warm-up, oracle port "
              f"of the reference modules (the reference's own torch CPU fp32 ops), {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": args.gpus,
            "steps": done, "warmup": warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": args.ref_batch, "parallelism": "cpu",
                       "sample": "bounded sample of the workload: same network and schedule, batch 64 per step"},
            "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if local == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from mcgm_b200 import _lib, cfg, models, ops, process_control, train_gan
    _lib.load()

    cfg.update(device=str(dev), data_name="CIFAR10", model_name="mcgan", control={"controller_rate": "0.5"})
    process_control()
    cfg["classes_size"] = 10
    torch.manual_seed(0)                      # identical weights / codebooks / SN vectors on every rank
    model = models.mcgan().to(dev)
    model.train(True)
    use_graph = not args.no_graph                    # --no-graph: eager step, all-reduce overlapped with backward
    opt = train_gan.make_optimizers(model, capturable=use_graph and world == 1)
    reducers = None
    if world > 1:
        reducers = {"generator": train_gan.GradAllReduce(model.generator, overlap=not use_graph),
                    "discriminator": train_gan.GradAllReduce(model.discriminator, overlap=not use_graph)}
    B = args.batch
    g = torch.Generator().manual_seed(1 + rank)            # per-rank data shard
    n_host = 4
    host_img = [(torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).pin_memory() for _ in range(n_host)]
    host_lab = [torch.randint(0, 10, (B,), generator=g).pin_memory() for _ in range(n_host)]
    dev_img = [t.to(dev) for t in host_img]
    dev_lab = [t.to(dev) for t in host_lab]
    host_loss = torch.zeros(3, dtype=torch.float32).pin_memory()
    torch.manual_seed(100 + rank)                          # z stream

    graphed = None
    if use_graph:
        graphed = (train_gan.GraphedTrainStep(model, opt, B) if world == 1
                   else train_gan.GraphedDDPTrainStep(model, opt, reducers, B))

    def run_step(batch):
        if graphed is not None:
            return graphed(batch)           # copies the batch into the graph's static buffers, replays
        return train_gan.train_step(model, opt, batch, reducers=reducers)

    def step_resident(i):
        return run_step({"img": dev_img[i % n_host], "label": dev_lab[i % n_host]})

    def step_e2e(i):
        if graphed is not None:             # pinned host -> static device buffers inside the call
            out = run_step({"img": host_img[i % n_host], "label": host_lab[i % n_host]})
        else:
            out = run_step({"img": host_img[i % n_host].to(dev, non_blocking=True),
                            "label": host_lab[i % n_host].to(dev, non_blocking=True)})
        host_loss.copy_(torch.stack([out["loss"], out["loss_D"], out["loss_G"]]).float(), non_blocking=True)
        return out

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _lib.launch_count
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = graphed.native_launches * steps if graphed is not None else _lib.launch_count - c0
        return float(ms), launches

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(2):                      # every rank (the step holds collectives); gives nvidia-smi
        step_resident(i)                    # the ~0.3 s it needs before its first sample
    torch.cuda.synchronize()
    sampler.rows.clear()
    ms, launches = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e(0)
    ms_e2e, _ = timed(step_e2e, args.steps)
    loss_host = [float(v) for v in host_loss]

    # per-kernel roofline: one extra instrumented iteration (CUDA events around each native call)
    roof = roof_gate = roof_wgrad = None
    if not args.no_kernel_profile:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "_src": "fallback"}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks.update(json.load(f)); peaks["_src"] = "measured"
        except Exception:
            pass
        # one untimed eager iteration first: it grows the caching allocator to the eager working set (the timed
        # runs above lived in the CUDA graph's private pool), so the measured pass below issues no cudaMalloc
        train_gan.train_step(model, opt, {"img": dev_img[0], "label": dev_lab[0]}, reducers=reducers)
        torch.cuda.synchronize()
        ops.PROFILE = {}
        # eager (un-graphed) iteration so that CUDA events can bracket each native call.  Eager issue is
        # host-bound (the GPU would idle between launches and the idle time would land inside the event
        # pairs of the small kernels), so the stream is first held back by a spin kernel long enough for the
        # host to enqueue the whole iteration: the events then see back-to-back device execution.
        torch.cuda._sleep(int(0.25 * 1.9e9))
        train_gan.train_step(model, opt, {"img": dev_img[0], "label": dev_lab[0]}, reducers=reducers)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None

        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                ncu_traffic = json.load(f)            # per-launch DRAM bytes of the dominant shapes (ncu --set full)
        except Exception:
            ncu_traffic = {}
        step_s = ms / args.steps * 1e-3

        def agg(tag, scale):
            """-> (launches, seconds, work, top shapes) of one kernel family in the profiled iteration."""
            recs = prof.get(tag, [])
            t = sum(r[0].elapsed_time(r[1]) for r in recs) * 1e-3
            shapes = {}
            for r in recs:
                e = shapes.setdefault(r[3], [0, 0.0, 0.0])
                e[0] += 1; e[1] += r[0].elapsed_time(r[1]) * 1e-3; e[2] += r[2]
            top = []
            for desc, (n, ts, w) in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:3]:
                tr = ncu_traffic.get(desc)
                top.append({"shape": desc, "launches": n, "avg_us": ts / n * 1e6, "achieved": w / ts / scale,
                            "share_of_step": ts / step_s,
                            "traffic": (tr["dram_read_bytes"] + tr["dram_write_bytes"]) if tr else None,
                            "ncu_tensor_pipe_active_pct": tr["tensor_pipe_active_pct"] if tr else None})
            return len(recs), t, sum(r[2] for r in recs), top

        def roof_obj(kernel, bound, tag, peak, unit, scale, src):
            n, t, work, top = agg(tag, scale)
            if not n:
                return None
            return {"kernel": kernel, "bound": bound, "achieved": work / t / scale, "peak": peak, "unit": unit,
                    "frac": work / t / scale / peak,
                    # DRAM bytes of ONE launch of the family's top shape (profiles/r01_traffic.json, ncu --set full)
                    "traffic": top[0]["traffic"] if top else None, "traffic_shape": top[0]["shape"] if top else None,
                    "launches": n, "avg_us": t / n * 1e6, "share_of_step": t / step_s, "peak_source": src,
                    "top_shapes": top}
        roof = roof_obj("conv_fwd_kernel (tcgen05 implicit GEMM; forward + dgrad launches)", "tensor", "conv",
                        peaks["bf16_tflops_sustained"], "TFLOP/s", 1e12, peaks["_src"] + " (cuBLAS bf16 sustained)")
        roof_wgrad = roof_obj("conv_wgrad_kernel (+split-K reduce, bias-grad)", "tensor", "wgrad",
                              peaks["bf16_tflops_sustained"], "TFLOP/s", 1e12, peaks["_src"] + " (cuBLAS bf16 sustained)")
        roof_gate = roof_obj("MC gate kernels (bn_partial/finalize, gate_fwd, gate_bwd_reduce/apply)", "hbm", "gate",
                             peaks["hbm_gbs"], "GB/s", 1e9, peaks["_src"] + " (copy)")

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, per, done = cpu_reference(64, 2, 1, budget_s=25.0)
        cpu_base = {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
                    "sample": f"{done} iteration(s) (5D+1G) at batch 64 (BASELINE configs[0]) after 1 warm-up "
                              f"({per:.2f} s/iteration), oracle port of the reference modules (torch CPU fp32)"}

    if rank == 0:
        per = ms / args.steps
        line = {"metric": METRIC, "value": B * world / (per * 1e-3), "unit": "img/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": B * world, "per_gpu_batch": B,
                           "parallelism": f"dp{world}", "schedule": "5D+1G per step",
                           "launch": (("one CUDA graph per iteration" if world == 1 else
                                       "CUDA graph per D/G forward+backward, NCCL all-reduce + Adam between graphs")
                                      if use_graph else "eager, bucketed all-reduce overlapped with backward"),
                           "l2": "no flush needed: each step streams >10 GB of activations (>> 126 MB L2)",
                           "kernel_profile": "one extra eager iteration after the timed region after an untimed one, behind a 250 ms spin kernel (no host-induced gaps), CUDA events per native call"},
                "e2e": {"value": B * world / (ms_e2e / args.steps * 1e-3), "unit": "img/s",
                        "h2d_bytes_per_step": B * 3 * 32 * 32 * 4 + B * 8, "d2h_bytes_per_step": 12,
                        "ms_per_step": ms_e2e / args.steps, "loss_D": loss_host[1], "loss_G": loss_host[2]},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_gate": roof_gate,
                "roofline_wgrad": roof_wgrad, "cpu_baseline": cpu_base}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
