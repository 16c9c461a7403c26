#!/usr/bin/env python
"""bench.py -- images/s of the FD-GAN training step (G + Fusion-D + VGG16 perceptual loss) at 256x256, batch 16 per GPU.

    python bench.py --gpus N --steps K --warmup W              # fdgan_b200 arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU arithmetic (oracle port) on host cores

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with the batch already resident in HBM; `e2e` = the same
through the public call (GANTrainer.step) with pinned-host inputs copied H2D and the loss scalars read D2H inside the
timed region.  `roofline` is measured live: every fdgan_b200 launch of the dominant kernel family is bracketed by CUDA
events on its stream during a profiled pass of the same steps (fdg_profile_*), algorithmic FLOPs come from the
descriptors.  `cpu_baseline` times oracle/fdgan_oracle.py (a port, the reference has no train.py) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec, FD-GAN train step (G+D+VGG16 perceptual), 256x256, batch 16 per GPU"
UNIT = "images/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


def ncu_family_profile(family):
    """Per-launch DRAM traffic / tensor-pipe activity of a kernel family from the committed ncu launch list of this same
    command (profiles/r01g_traffic.json; numbers taken under ncu are never bench values, they only annotate the roofline)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01g_traffic.json")) as f:
            return json.load(f)["families"].get(family)
    except Exception:
        return None


def synth_pair(index: int, size: int):
    """J ~ U[0,1), t ~ U[0.3,0.9], A ~ U[0.7,1.0] scalars, I = J t + A (1 - t); seed 1234 + index."""
    g = torch.Generator().manual_seed(1234 + index)
    J = torch.rand((3, size, size), generator=g)
    t = 0.3 + 0.6 * torch.rand((), generator=g)
    A = 0.7 + 0.3 * torch.rand((), generator=g)
    return J * t + A * (1 - t), J


def make_batches(n_batches, per_gpu, size, rank, pin):
    out = []
    for b in range(n_batches):
        hz, cl = [], []
        for i in range(per_gpu):
            h, c = synth_pair((b * 10007 + rank * per_gpu + i) % 100000, size)
            hz.append(h)
            cl.append(c)
        hz, cl = torch.stack(hz), torch.stack(cl)
        if pin:
            hz, cl = hz.pin_memory(), cl.pin_memory()
        out.append((hz, cl))
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(B, world, size):
    """`config` of the JSON line: BASELINE.json configs[2]; shared by both arms so that the driver compares like with like."""
    return {"workload": "G+D+VGG16 perceptual train step (SURVEY 3.3 reconstruction): FDGAN fwd+bwd, D(9,36) x3 fwd / x2 wgrad / x1 dgrad, "
                        "freq decomposition fwd+bwd, Vgg16 fwd x2 + dgrad, Adam x2",
            "per_gpu_batch": B, "global_batch": B * world, "image": "%dx%d" % (size, size),
            "loss": "L1 + 0.5*MSE(vgg relu2_2, relu4_3) + 0.01*BCE adversarial; D: BCE real/fake",
            "parallelism": "dp%d (batch shard, one NCCL all-reduce per network per step)" % world,
            "l2": "inputs larger than L2: each step streams several GB of activations (>> 126 MB L2); no explicit flush"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference arithmetic)
# ------------------------------------------------------------------------------------------------------------------


def cpu_step_images_per_s(sample_batch, size, steps, warmup):
    from oracle import fdgan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g_sd, d_sd, v_sd = O.make_fdgan_state(0), O.make_d_state(9, 36, 1), O.make_vgg_state(2)
    sg, sdd = {}, {}
    pairs = [synth_pair(i, size) for i in range(sample_batch)]
    hazy, clean = torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])
    for _ in range(warmup):
        O.train_step(g_sd, d_sd, v_sd, hazy, clean, sg, sdd)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(g_sd, d_sd, v_sd, hazy, clean, sg, sdd)
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 2
    steps, warmup = max(1, args.steps), max(0, args.warmup)      # exactly K timed steps of the bounded sample (~2 s each on 16 cores)
    ips, s_per_step, cores = cpu_step_images_per_s(sample, args.size, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.batch, args.gpus, args.size),      # the fdgan_b200 arm's config, key for key
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d-image batches, %d timed steps of the full step (oracle/fdgan_oracle.py:train_step)" % (sample, steps)},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# fdgan_b200 arm
# ------------------------------------------------------------------------------------------------------------------


def run_ours(args):
    import fdgan_b200
    from fdgan_b200 import _lib as L
    from fdgan_b200 import dist as fdist
    from fdgan_b200.train import GANTrainer
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device for the fdgan_b200 arm (there is no CPU fallback)"
    rank, local_rank, world = fdist.init_process_group()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(0)
    G, D, V = fdgan_b200.FDGAN(), fdgan_b200.D(9, 36), fdgan_b200.Vgg16()
    torch.manual_seed(2)
    G, D, V = G.to(dev).train(), D.to(dev).train(), V.to(dev)
    for p in V.parameters():
        p.requires_grad_(False)
    tr = GANTrainer(G, D, V)
    B, size = args.batch, args.size
    host = make_batches(2, B, size, rank, pin=True)
    resident = [(h.to(dev), c.to(dev)) for h, c in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- warm-up (also pages in kernels / allocator pools)
    for i in range(max(3, args.warmup)):
        h, c = resident[i % len(resident)]
        tr.step(h, c)
    barrier()

    # ---- value: inputs resident in HBM
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.launch_count()
    with ClockSampler(local_rank) as clk:
        barrier()
        if args.quick:
            torch.cuda.profiler.start()      # ncu --profile-from-start off captures exactly the timed steps
        ev0.record()
        for i in range(args.steps):
            h, c = resident[i % len(resident)]
            tr.step(h, c, sync_losses=False)
        ev1.record()
        barrier()
        if args.quick:
            torch.cuda.profiler.stop()
    launches = L.launch_count() - n0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    clocks = clk.summary()

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": value, "unit": UNIT, "ms_per_step": ms_step, "gpu_launches": int(launches),
                              "steps": args.steps, "clocks": clocks}))
        return 0
    # ---- e2e: pinned host inputs -> H2D, step, loss scalars D2H, all inside the timed region
    barrier()
    ev0.record()
    for i in range(args.steps):
        h, c = host[i % len(host)]
        tr.step(h.to(dev, non_blocking=True), c.to(dev, non_blocking=True), sync_losses=True)
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1))
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    last = dict(tr.last)

    # ---- live roofline of the dominant kernel family (profiled pass over the same steps; events on the launch stream)
    pk = peaks()
    roof = None
    fam = {}
    if rank == 0 or world == 1:
        pass
    L.profile_enable(True)
    prof_steps = min(args.steps, 3)
    for i in range(prof_steps):
        h, c = resident[i % len(resident)]
        tr.step(h, c, sync_losses=False)
    torch.cuda.synchronize()
    L.profile_enable(False)
    fam = L.profile_collect()
    tot_ms = sum(v["ms"] for v in fam.values())
    dom = max(fam, key=lambda k: fam[k]["ms"])
    d = fam[dom]
    if d["ms"] > 0:
        if dom in ("conv_simt_f32", "conv_tcgen05", "wgrad"):
            achieved = d["flops"] / (d["ms"] / 1e3) / 1e12
            prof = ncu_family_profile(dom)
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tf_sustained"], "traffic": prof["dram_bytes_per_launch"] if prof else None, "kernel": dom,
                    "traffic_source": "profiles/r01g_traffic.json: mean dram__bytes_read+write per launch of this family (ncu, same command)" if prof else None,
                    "tensor_pipe_active_pct_ncu": prof["tensor_pipe_active_pct"] if prof else None,
                    "note": "fp32 operands run as bf16 hi/lo splits: 3 algorithmic bf16 passes per MAC (issued as 2 MMAs of width 2N and N), "
                            "so the algorithmic ceiling is 1/3 of the bf16 peak; achieved counts each MAC once",
                    "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                    "share_of_profiled_kernel_time": d["ms"] / tot_ms, "launches_per_step": d["launches"] / prof_steps,
                    "avg_launch_ms": d["ms"] / max(1, d["launches"]),
                    "flops_per_launch": d["flops"] / max(1, d["launches"])}
        else:
            achieved = d["bytes"] / (d["ms"] / 1e3) / 1e9
            prof = ncu_family_profile(dom)
            roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                    "traffic": prof["dram_bytes_per_launch"] if prof else None, "kernel": dom, "peak_source": pk["src"],
                    "share_of_profiled_kernel_time": d["ms"] / tot_ms}
    families = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                    "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 else 0.0,
                    "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 else 0.0} for k, v in fam.items() if v["launches"]}

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, s_per, cores = cpu_step_images_per_s(2, size, 2, 1)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2-image batches, 2 timed steps (+1 warm-up) of the full step, oracle/fdgan_oracle.py:train_step, %.1f s/step" % s_per}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world, size),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * B * 3 * size * size * 4, "d2h_bytes_per_step": 40,      # GANTrainer.loss_buf: five fp64 loss terms
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof,
            "kernel_families": families,
            "cpu_baseline": cpu,
            "losses_last_step": last,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fdgan_b200", choices=["fdgan_b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU (weak scaling)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="value leg only (used under ncu): no e2e / profiled / CPU passes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
