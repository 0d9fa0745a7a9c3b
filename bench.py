"""bench.py — training throughput (images/s) of the MC-gated models on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3                       # headline: MCGAN CIFAR10-shaped
    python bench.py --config {cifar10,omniglot,mcvae,mcglow}             # BASELINE.json configs[1], [2], [3], [4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host cores

Default workload (BASELINE.json `metric`): a "step" is ONE reference training iteration
(src/train_gan.py:139-176: 5 D updates + 1 G update, hinge loss, Adam) of the CIFAR10-shaped MCGAN
(32x32x3, 10 modes, controller rate 0.5, G widths 256x4, D widths 128x4, latent 128) at batch 512 per
GPU, bf16 activations with fp32 accumulation, synthetic data, random-init weights (seed 0).  One JSON
line is printed by rank 0.  See DESIGN.md "Measurement" for what each key means.
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

CONFIGS = {
    # name: kind, data_name, model_name, classes, per-GPU batch, CPU-arm batch, metric, workload
    "cifar10": dict(kind="gan", data="CIFAR10", model="mcgan", classes=10, batch=512, ref_batch=64, strides=[2, 1, 1],
                    metric="MCGAN G+D train imgs/s",
                    workload="MCGAN CIFAR10-shaped 32x32x3, 10 modes, control 0.5, one iteration = 5 D + 1 G updates "
                             "(train_gan.py:139-176)"),
    "omniglot": dict(kind="gan", data="Omniglot", model="mcgan", classes=1623, batch=512, ref_batch=16, strides=[2, 2, 1],
                     metric="MCGAN G+D train imgs/s (Omniglot-shaped, 1623 modes)",
                     workload="MCGAN Omniglot-shaped 32x32x1, 1623 modes, control 0.5, G widths 512/256/128/64, D widths "
                              "64/128/256/512, one iteration = 5 D + 1 G updates (train_gan.py:139-176)"),
    "mcvae": dict(kind="vae", data="COIL100", model="mcvae", classes=100, batch=512, ref_batch=64,
                  metric="MCVAE train imgs/s",
                  workload="MCVAE COIL100-shaped 32x32x3, 100 modes, control 0.5, widths 64/128/256, latent 128, 2 ResBlocks; "
                           "one step = encoder+decoder forward, backward, clip_grad_norm_(1), Adam (train_vae.py:106-111)"),
    "mcglow": dict(kind="glow", data="CIFAR10", model="mcglow", classes=10, batch=128, ref_batch=8,
                   metric="MCGlow train imgs/s",
                   workload="MCGlow CIFAR10-shaped 32x32x3, K=16, L=3, hidden 512, affine coupling, LU invertible conv, "
                            "10 modes; one step = forward, backward, clip_grad_norm_(1), Adam (train_glow.py:116-121)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cifar10", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0: the config's default)")
    ap.add_argument("--ref-batch", type=int, default=0, help="images per step of the CPU reference arm (0: the config's default)")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="time budget of the CPU reference arm's timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    ap.add_argument("--no-torch-eager", action="store_true", help="skip the informational PyTorch-eager-on-this-GPU leg")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--ddp", default="auto", choices=["auto", "ingraph", "segmented"],
                    help="N>1 launch form: NCCL all-reduce captured inside the CUDA graph, or between graph segments")
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
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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


# ----------------------------------------------------------------------------- model construction (both arms)
def configure(C, device):
    """Fill the global cfg the way a user of the reference would (config.yml + process_control)."""
    from mcgm_b200 import cfg, process_control
    cfg.update(device=str(device), data_name=C["data"], model_name=C["model"], control={"controller_rate": "0.5"})
    process_control()
    cfg["classes_size"] = C["classes"]
    return cfg


def build_model(C, device):
    import numpy as np
    import torch
    from mcgm_b200 import models
    configure(C, device)
    torch.manual_seed(0)                      # identical weights / codebooks / SN vectors on every rank
    np.random.seed(0)                         # Glow's QR initialisation (src/models/mcglow.py:79)
    return getattr(models, C["model"])()


def synth_batch(C, batch, gen):
    import torch
    from mcgm_b200 import cfg
    img = torch.rand(batch, *cfg["data_shape"], generator=gen) * 2 - 1
    label = torch.randint(0, C["classes"], (batch,), generator=gen)
    return img, label


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def oracle_stepper(C, batch, device="cpu"):
    """-> a callable running ONE reference step (oracle port: the reference's own torch ops) on `device`."""
    import torch
    model = build_model(C, "cpu")
    g = torch.Generator().manual_seed(1)
    img, label = synth_batch(C, batch, g)
    img, label = img.to(device), label.to(device)
    if C["kind"] == "gan":
        from oracle import mcgan_oracle as O
        sd = O.make_trainable({k: v.to(device) for k, v in model.state_dict().items()})
        opt = O.make_adams(sd)

        def step():
            zs = [torch.randn(batch, 128, generator=g).to(device) for _ in range(6)]
            return O.train_iteration(sd, opt, img, label, zs, C["classes"], C["strides"])
        return step
    if C["kind"] == "vae":
        from oracle import mcvae_oracle as V
        sd = V.make_trainable({k: v.to(device) for k, v in model.state_dict().items()})
        opt = V.make_adam(sd)

        def step():
            eps = torch.randn(batch, 128, generator=g).to(device)
            return V.train_step(sd, opt, img, label, C["classes"], eps)["loss"]
        return step
    from oracle import mcglow_oracle as G
    sd = G.make_trainable({k: v.to(device) for k, v in model.state_dict().items()})
    with torch.no_grad():                     # data-dependent ActNorm init (src/train_glow.py:60-67)
        G.forward(img, label, sd, C["classes"], torch.rand(img.shape, generator=g).to(device))
    opt = G.make_adam(sd)

    def step():
        noise = torch.rand(img.shape, generator=g).to(device)
        return G.train_step(sd, opt, img, label, C["classes"], noise)["loss"]
    return step


def cpu_reference(C, batch, steps, warmup, threads=None, budget_s=None):
    """The reference algorithm on the host cores: up to `steps` steps at `batch` images (fewer if `budget_s`
    would be exceeded, judged from the timed steps so far).  -> (img/s, cores, s/step, steps timed)."""
    import torch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    step = oracle_stepper(C, batch)
    times = []
    i = 0
    while i < warmup + steps:
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            if budget_s is not None and sum(times) + dt > budget_s:
                break
        i += 1
    per = sum(times) / len(times)
    return batch / per, torch.get_num_threads(), per, len(times)


def torch_eager_gpu(C, batch, dev):
    """Informational: the SAME reference ops (oracle functions, i.e. ATen / cuDNN / cuBLAS eager kernels, NCHW)
    on this GPU — the like-for-like GPU context number BASELINE.md §3 asks for.  TF32 and bf16 autocast."""
    import torch
    out = {"what": "oracle port of the reference step on CUDA tensors (ATen/cuDNN/cuBLAS eager, NCHW), "
                   f"batch {batch}, 3 timed steps after 2 warm-up"}
    for mode in ("tf32", "bf16_autocast"):
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cudnn.benchmark = True          # src/train_gan.py:13
            step = oracle_stepper(C, batch, dev)
            import contextlib
            ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if mode == "bf16_autocast" else contextlib.nullcontext
            for _ in range(2):
                with ctx():
                    step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                with ctx():
                    step()
            torch.cuda.synchronize()
            per = (time.perf_counter() - t0) / 3
            out[mode] = {"value": batch / per, "unit": "img/s", "ms_per_step": per * 1e3}
        except Exception as e:                              # informational leg: never fails the bench
            out[mode] = {"error": f"{type(e).__name__}: {e}"[:200]}
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.benchmark = False
    torch.cuda.empty_cache()
    return out


def run_reference(args, C):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rb = args.ref_batch or C["ref_batch"]
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    value, cores, per, done = cpu_reference(C, rb, steps, warmup, budget_s=args.ref_seconds)
    sample = (f"{done} step(s) at batch {rb} after {warmup} warm-up, oracle port of the reference modules "
              f"(the reference's own torch CPU fp32 ops), {cores} threads")
    line = {"impl": "reference", "metric": C["metric"], "value": value, "unit": "img/s", "n_gpus": args.gpus,
            "steps": done, "warmup": warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C["workload"], "global_batch": rb, "parallelism": "cpu",
                       "sample": f"bounded sample of the workload: same network and schedule, batch {rb} per step"},
            "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- native arm
class Workload:
    """One config on the B200 kernels: model, optimiser(s), the step as the user calls it, an eager step for
    the per-kernel profile."""

    def __init__(self, C, args, dev, rank, world):
        import torch
        from mcgm_b200 import train_gan, train_glow, train_vae
        self.C, self.dev, self.world = C, dev, world
        self.B = args.batch or C["batch"]
        self.model = build_model(C, dev).to(dev)
        self.model.train(True)
        self.use_graph = not args.no_graph
        self.reducers = None
        self.launch_form = ""
        g = torch.Generator().manual_seed(1 + rank)            # per-rank data shard
        if C["kind"] == "gan":
            self.opt = train_gan.make_optimizers(self.model, capturable=self.use_graph and world == 1)
            if world > 1:
                self.reducers = {k: train_gan.GradAllReduce(getattr(self.model, k), overlap=not self.use_graph)
                                 for k in ("generator", "discriminator")}
            torch.manual_seed(100 + rank)                      # z stream
            self.graphed = None
            if self.use_graph:
                if world == 1:
                    self.graphed = train_gan.GraphedTrainStep(self.model, self.opt, self.B)
                    self.launch_form = "one CUDA graph per iteration"
                else:
                    self.graphed = train_gan.make_ddp_step(self.model, self.opt, self.reducers, self.B, mode=args.ddp)
                    self.launch_form = self.graphed.launch_form
            else:
                self.launch_form = "eager, bucketed all-reduce overlapped with backward"
            self.loss_keys = ("loss", "loss_D", "loss_G")
        else:
            if world > 1:
                raise SystemExit(f"--config {C['model']}: the data-parallel bench covers the MCGAN configs; run this one at --gpus 1")
            self.opt = train_vae.make_optimizer(self.model)
            img, label = synth_batch(C, self.B, g)
            ex = {"img": img.to(dev), "label": label.to(dev)}
            if C["kind"] == "glow":
                train_glow.init_actnorm(self.model, ex)        # src/train_glow.py:60-67 (one batch here)
            self.graphed = train_vae.GraphedTrainStep(self.model, self.opt, ex) if self.use_graph else None
            self.launch_form = "one CUDA graph per step" if self.use_graph else "eager"
            self.loss_keys = ("loss",)
        self.gen = g

    def step(self, batch):
        """The call a user makes: {'img','label'} on the host (pinned) or the device -> dict of device losses."""
        import torch
        from mcgm_b200 import train_gan, train_vae
        if self.graphed is not None:
            return self.graphed(batch)           # copies the batch into the graph's static buffers, replays
        b = {k: v.to(self.dev, non_blocking=True) for k, v in batch.items()}
        if self.C["kind"] == "gan":
            return train_gan.train_step(self.model, self.opt, b, reducers=self.reducers)
        return train_vae.train_step(self.model, self.opt, b)

    def eager_step(self, batch):
        from mcgm_b200 import train_gan, train_vae
        if self.C["kind"] == "gan":
            return train_gan.train_step(self.model, self.opt, batch, reducers=self.reducers)
        out = train_vae.train_step(self.model, self.opt, batch)
        return {"loss": out["loss"].detach()}

    def losses(self, out):
        import torch
        return torch.stack([out[k].detach().float().reshape(()) for k in self.loss_keys])

    @property
    def native_launches(self):
        return self.graphed.native_launches if self.graphed is not None else None


def run_native(args, C):
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
    from mcgm_b200 import _lib, data, ops
    _lib.load()

    W = Workload(C, args, dev, rank, world)
    B = W.B
    n_host = 4
    host = []
    for _ in range(n_host):
        img, lab = synth_batch(C, B, W.gen)
        host.append({"img": img.pin_memory(), "label": lab.pin_memory()})
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    n_loss = len(W.loss_keys)
    host_loss = torch.zeros(n_loss, dtype=torch.float32).pin_memory()

    def step_resident(i):
        return W.step(resident[i % n_host])

    def timed_device(fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""
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
        launches = W.native_launches * steps if W.native_launches is not None else _lib.launch_count - c0
        return float(ms), launches

    warm = max(args.warmup, 3)
    for i in range(warm):
        step_resident(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(2):                      # every rank (the step holds collectives); gives nvidia-smi
        step_resident(i)                    # the ~0.3 s it needs before its first sample
    torch.cuda.synchronize()
    sampler.rows.clear()
    ms, launches = timed_device(step_resident, args.steps)
    if ms < 600.0:                          # a region shorter than a few sampling periods (MCVAE: 50 ms): keep the same
        for i in range(int(600.0 / max(ms / args.steps, 1e-3)) + 1):     # load running, untimed, so that the clocks line has samples
            step_resident(i)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: a host producer thread collates each batch from a host-resident uint8 image set into pinned
    # staging buffers, the feeder copies it to the device on a copy stream (H2D inside the timed region), the step
    # runs through the public call, the losses are read back (D2H) every step.  Timed with a host wall clock.
    dataset = data.SyntheticImages(8 * B, [int(s) for s in resident[0]["img"].shape[1:]], C["classes"], seed=5 + rank)
    idx_gen = torch.Generator().manual_seed(9 + rank)

    def loader(n_steps):
        for _ in range(n_steps):
            yield dataset.batch(torch.randint(0, len(dataset), (B,), generator=idx_gen))

    feeder = data.DeviceFeeder(None, dev, depth=2)

    def run_e2e(n_steps):
        out = None
        for batch in feeder.iterate(loader(n_steps)):
            out = W.step(batch)
            host_loss.copy_(W.losses(out), non_blocking=True)

    run_e2e(3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize()
    wall = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    ms_e2e = float(wall) * 1e3
    loss_host = [float(v) for v in host_loss]

    # ---- per-kernel roofline: one extra instrumented iteration (CUDA events around each native call)
    roofs = {}
    if not args.no_kernel_profile:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "_src": "fallback"}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks.update(json.load(f)); peaks["_src"] = "measured"
        except Exception:
            pass
        # one untimed eager iteration first: it grows the caching allocator to the eager working set (the timed
        # runs above lived in the CUDA graph's private pool), so the measured pass below issues no cudaMalloc
        W.eager_step(resident[0])
        torch.cuda.synchronize()
        ops.PROFILE = {}
        # eager (un-graphed) iteration so that CUDA events can bracket each native call.  Eager issue is
        # host-bound (the GPU would idle between launches and the idle time would land inside the event
        # pairs of the small kernels), so the stream is first held back by a spin kernel long enough for the
        # host to enqueue the whole iteration: the events then see back-to-back device execution.
        torch.cuda._sleep(int((0.25 if C["kind"] != "glow" else 0.6) * 1.9e9))
        W.eager_step(resident[0])
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None

        ncu_traffic = {}
        for name in ("r01_traffic.json", "r02_traffic.json"):   # per-launch DRAM bytes of the dominant shapes (ncu --set full)
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    ncu_traffic.update(json.load(f))
            except Exception:
                pass
        step_s = ms / args.steps * 1e-3

        def roof_obj(kernel, bound, tag, peak, unit, scale, src):
            recs = prof.get(tag, [])
            if not recs:
                return None
            t = sum(r[0].elapsed_time(r[1]) for r in recs) * 1e-3
            work, ref_work = sum(r[2] for r in recs), sum(r[4] for r in recs)
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
            return {"kernel": kernel, "bound": bound, "achieved": work / t / scale, "peak": peak, "unit": unit,
                    # `frac`: EXECUTED work (the MACs the tensor cores run / the bytes the op must move) over time over peak.
                    # `frac_reference_op`: the same launches booked at the reference op's count (Up2->Conv3x3 = 9 taps per
                    # output where the kernel executes 4) — how fast the REFERENCE's arithmetic is being retired; not a roofline.
                    "frac": work / t / scale / peak, "frac_reference_op": ref_work / t / scale / peak,
                    # DRAM bytes of ONE launch of the family's top shape (profiles/r0x_traffic.json, ncu --set full)
                    "traffic": top[0]["traffic"] if top else None, "traffic_shape": top[0]["shape"] if top else None,
                    "launches": len(recs), "avg_us": t / len(recs) * 1e6, "share_of_step": t / step_s, "peak_source": src,
                    "top_shapes": top}
        tsrc, hsrc = peaks["_src"] + " (cuBLAS bf16 sustained)", peaks["_src"] + " (copy)"
        roofs["roofline"] = roof_obj("conv_fwd_kernel (tcgen05 implicit GEMM; forward + dgrad launches)", "tensor", "conv",
                                     peaks["bf16_tflops_sustained"], "TFLOP/s", 1e12, tsrc)
        roofs["roofline_wgrad"] = roof_obj("conv_wgrad_kernel (+split-K reduce, bias-grad)", "tensor", "wgrad",
                                           peaks["bf16_tflops_sustained"], "TFLOP/s", 1e12, tsrc)
        roofs["roofline_gate"] = roof_obj("MC gate kernels (bn_stats, gate_fwd, gate_bwd_reduce/apply)", "hbm", "gate",
                                          peaks["hbm_gbs"], "GB/s", 1e9, hsrc)
        roofs["roofline_thin"] = roof_obj("RGB-side (3-channel) convs: thin_conv_fwd / thin_conv_wgrad / im2col+GEMM+col2im "
                                          "(bandwidth-bound, SURVEY.md §8d)", "hbm", "thin", peaks["hbm_gbs"], "GB/s", 1e9, hsrc)

    cpu_base = eager_gpu = None
    if rank == 0 and world == 1 and not args.no_torch_eager:
        eager_gpu = torch_eager_gpu(C, B if C["kind"] != "glow" else min(B, 64), dev)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rb = args.ref_batch or C["ref_batch"]
        v, cores, per, done = cpu_reference(C, rb, 2, 1, budget_s=25.0)
        cpu_base = {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
                    "sample": f"{done} step(s) at batch {rb} after 1 warm-up ({per:.2f} s/step), oracle port of the "
                              f"reference modules (torch CPU fp32)"}

    if rank == 0:
        per = ms / args.steps
        bytes_in = sum(v.numel() * v.element_size() for v in host[0].values())
        line = {"metric": C["metric"], "value": B * world / (per * 1e-3), "unit": "img/s", "n_gpus": world,
                "steps": args.steps, "warmup": warm, "ms_per_step": per,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": C["workload"], "name": args.config, "global_batch": B * world, "per_gpu_batch": B,
                           "parallelism": f"dp{world}", "launch": W.launch_form,
                           "l2": "no flush needed: each step streams far more activation bytes than the 126 MB L2",
                           "e2e": "host producer thread (uint8 image set -> float [-1,1] batches into pinned staging), H2D on a copy "
                                  "stream, step through the public call, losses read back every step; host wall clock",
                           "kernel_profile": "one extra eager step after the timed region (after an untimed one), behind a spin "
                                             "kernel (no host-induced gaps), CUDA events per native call"},
                "e2e": {"value": B * world / (ms_e2e / args.steps * 1e-3), "unit": "img/s",
                        "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 4 * n_loss,
                        "ms_per_step": ms_e2e / args.steps, "clock": "time.perf_counter",
                        "losses": dict(zip(W.loss_keys, loss_host))},
                "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu_base, "torch_eager_b200": eager_gpu}
        line.update(roofs)
        if "roofline" not in line:
            line["roofline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        from mcgm_b200 import train_gan
        train_gan.shutdown_process_group(W.graphed)


if __name__ == "__main__":
    a = parse()
    conf = CONFIGS[a.config]
    if a.impl == "reference":
        run_reference(a, conf)
    else:
        run_native(a, conf)
