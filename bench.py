#!/usr/bin/env python
"""bench.py -- images/s of the FD-GAN training step (G + Fusion-D + VGG16 perceptual loss) at 256x256, batch 16 per GPU.

    python bench.py --gpus N --steps K --warmup W              # fdgan_b200 arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own modules on the host cores (CPU path)

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with the batch already resident in HBM; `e2e` = the same
through the public call (GANTrainer.step) with pinned-host inputs copied H2D and the loss scalars read D2H inside the
timed region.  `roofline` is measured live: every fdgan_b200 launch of the dominant kernel family is bracketed by CUDA
events on its stream during a profiled pass of the same steps (fdg_profile_*), algorithmic FLOPs come from the
descriptors.  `cpu_baseline` / `--impl reference` time the step composed from the reference's OWN modules (oracle/ref_step.py
over baseline/_ref, placed there by oracle/vendor_ref.py) on the host cores, on a bounded sample; `gpu_baseline` runs the same
reference modules through eager torch / cuDNN on the same B200 (the >= 6x denominator of BASELINE.json), outside the timed
region; `secondary` reports configs[1] (B=1 step) and configs[4] (1280x720 inference, B=4); `dp_check` (N > 1) proves that
the replicas stayed bit-identical and that the all-reduced gradient equals a single-process computation of the same shards.
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


TRAFFIC_FILE = os.path.join("profiles", "r02_traffic.json")


def kernel_source_sha():
    """sha256 over the kernel sources the library is built from (csrc/*.cu, *.cuh, include/*.h): ties an ncu launch list to a build."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "fdgan_b200", "csrc", "*.cu*")) + glob.glob(os.path.join(ROOT, "include", "*.h"))):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_family_profile(family):
    """Per-launch DRAM traffic / tensor-pipe activity of a kernel family from the committed ncu launch list of this same
    command (numbers taken under ncu are never bench values, they only annotate the roofline).  The file records the hash of
    the kernel sources it was captured on; a file from another build is refused (traffic = null, reason stated)."""
    try:
        with open(os.path.join(ROOT, TRAFFIC_FILE)) as f:
            js = json.load(f)
    except Exception:
        return None, "no " + TRAFFIC_FILE
    if js.get("src_sha") != kernel_source_sha():
        return None, "%s was captured on kernel sources %s, this build is %s: refused as stale" % (TRAFFIC_FILE, js.get("src_sha"), kernel_source_sha())
    return js["families"].get(family), "%s: mean dram__bytes_read+write per launch of this family (ncu launch list of this command, same kernel sources %s)" % (TRAFFIC_FILE, js["src_sha"])


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


def reference_stepper(device):
    """(callable(hazy, clean), kind): the step composed from the reference's own modules (baseline/_ref or /root/reference,
    kind "reference"); if neither is present, the oracle's restatement of the same arithmetic (kind "port")."""
    import warnings
    warnings.filterwarnings("ignore", message=".*upsample_nearest.*")
    from oracle import ref_step
    if ref_step.available():
        rs = ref_step.RefStep(device)
        return rs.step, "reference"
    from oracle import fdgan_oracle as O
    dev = torch.device(device)
    g_sd, d_sd, v_sd = (type(sd)((k, v.to(dev)) for k, v in sd.items()) for sd in (O.make_fdgan_state(0), O.make_d_state(9, 36, 1), O.make_vgg_state(2)))
    sg, sdd = {}, {}
    return (lambda hazy, clean: O.train_step(g_sd, d_sd, v_sd, hazy, clean, sg, sdd)), "port"


def cpu_step_images_per_s(sample_batch, size, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind = reference_stepper("cpu")
    pairs = [synth_pair(i, size) for i in range(sample_batch)]
    hazy, clean = torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])
    for _ in range(warmup):
        step(hazy, clean)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(hazy, clean)
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps, torch.get_num_threads(), kind


def gpu_baseline(B, size, dev, steps=3, warmup=2):
    """The reference's torch / cuDNN path on the same B200, same batch, same synthetic pairs (demo.py:11-12 sets
    cudnn.benchmark = True).  Strict fp32 (TF32 off) is the reference's arithmetic; TF32-allowed is torch's cuDNN default and does
    not hold the 1e-3 parity bar.  Timed with CUDA events, outside bench.py's timed region."""
    out = {"what": "eager torch %s / cuDNN %s, step composed from the reference's own FDGAN / D / Vgg16 modules + torch.optim.Adam "
                   "(oracle/ref_step.py), cudnn.benchmark=True, batch %d, %dx%d" % (torch.__version__, torch.backends.cudnn.version(), B, size, size),
           "unit": UNIT}
    pairs = [synth_pair(i, size) for i in range(B)]
    hazy, clean = torch.stack([p[0] for p in pairs]).to(dev), torch.stack([p[1] for p in pairs]).to(dev)
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("strict_fp32", False), ("tf32", True)):
            torch.backends.cudnn.benchmark = True
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            step, kind = reference_stepper(dev)
            for _ in range(warmup):
                step(hazy, clean)
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step(hazy, clean)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": 1e3 * B / ms, "ms_per_step": ms, "steps": steps, "warmup": warmup, "kind": kind,
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
            del step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.ref_sample      # bounded sample: K timed steps of a `sample`-image batch (a 16-image CPU step costs ~15 s and ~60 GB)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    ips, s_per_step, cores, kind = cpu_step_images_per_s(sample, args.size, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.batch, args.gpus, args.size),      # the fdgan_b200 arm's config, key for key
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample_batch": sample,
                         "sample": "each timed step is the full G+D+VGG step on a %d-image batch (NOT the %d of config.per_gpu_batch: BatchNorm sees %d "
                                   "images), %d timed steps; %s" % (sample, args.batch, sample, steps,
                                   "reference modules from baseline/_ref via oracle/ref_step.py" if kind == "reference" else "oracle/fdgan_oracle.py:train_step")},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "sample_batch": sample,
    }
    print(json.dumps(line))
    return 0


def dp_check(tr, world, rank, dev, size):
    """Run after the timed region at N > 1.  (i) every rank's flat parameter buffers must equal rank 0's bit for bit (max |diff| = 0
    after all the timed steps); (ii) one extra gradient-only step on a fixed 2*N-image batch: the all-reduced, rank-averaged flat
    gradients are compared on rank 0 with a single-process computation of the same N shards (per-shard BatchNorm statistics, as
    nn.DataParallel / DDP have them, demo.py:89)."""
    import torch.distributed as dist
    div = 0.0
    for st in (tr.sG, tr.sD):
        ref = st.flat.clone()
        dist.broadcast(ref, src=0)
        dmax = (st.flat - ref).abs().max().reshape(1)
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        div = max(div, float(dmax.item()))
    shards = []
    for r in range(world):
        prs = [synth_pair(50000 + 2 * r + i, size) for i in range(2)]
        shards.append((torch.stack([p[0] for p in prs]).to(dev), torch.stack([p[1] for p in prs]).to(dev)))
    h, c = shards[rank]
    tr.step(h, c, sync_losses=False, apply=False, allreduce=True)
    torch.cuda.synchronize()
    got = [tr.sG.grad.clone().div_(world), tr.sD.grad.clone().div_(world)]
    out = {"param_divergence": div, "shard_batch": 2}
    if rank == 0:
        want = [torch.zeros_like(got[0]), torch.zeros_like(got[1])]
        for r in range(world):
            tr.step(shards[r][0], shards[r][1], sync_losses=False, apply=False, allreduce=False)
            want[0] += tr.sG.grad
            want[1] += tr.sD.grad
        for k, g, w in (("grad_rel_l2_G", got[0], want[0].div_(world)), ("grad_rel_l2_D", got[1], want[1].div_(world))):
            out[k] = float((g - w).norm() / w.norm())
        out["grad_rel_l2"] = max(out["grad_rel_l2_G"], out["grad_rel_l2_D"])
        out["note"] = ("rank-0 recomputation of every shard against the NCCL result; not bit-exact because weight-gradient kernels accumulate "
                       "split-K partials with atomics (run-to-run order), expected ~1e-6")
    dist.barrier()
    return out


def secondary_configs(tr, G, world, rank, dev, size, pk, max_over_ranks, barrier):
    """configs[1]: G+D(+VGG) step at batch 1 (the reference's default batch, demo.py:33-34), eager and replayed from a CUDA graph
    (N = 1 only: the captured step has no all-reduce).  configs[4]: FDGAN forward at 1280x720, batch 4 per GPU, train-mode BatchNorm
    (README.md:38), every rank an independent replica -> whole-job images/s."""
    from fdgan_b200 import _lib as L
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world == 1:
        h1, c1 = (t.to(dev) for t in make_batches(1, 1, size, 0, pin=False)[0])
        for _ in range(3):
            tr.step(h1, c1, sync_losses=False)
        torch.cuda.synchronize()
        n0 = L.launch_count()
        ev0.record()
        for _ in range(20):
            tr.step(h1, c1, sync_losses=False)
        ev1.record()
        torch.cuda.synchronize()
        ms_eager = ev0.elapsed_time(ev1) / 20
        launches = (L.launch_count() - n0) / 20
        for _ in range(3):
            tr.step_graphed(h1, c1, sync_losses=False)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(20):
            tr.step_graphed(h1, c1, sync_losses=False)
        ev1.record()
        torch.cuda.synchronize()
        ms_graph = ev0.elapsed_time(ev1) / 20
        out["configs[1]_b1_step_256"] = {"ms_per_step_eager": ms_eager, "images_per_s_eager": 1e3 / ms_eager, "ms_per_step_graphed": ms_graph,
                                        "images_per_s_graphed": 1e3 / ms_graph, "launches_per_step": launches, "steps": 20, "warmup": 3,
                                        "workload": "same step as the headline (G+D+VGG16 perceptual) at batch 1, 256x256"}
    # ---- 1280x720 inference, batch 4 per GPU
    Bi = 4
    g = torch.Generator().manual_seed(777 + rank)
    x = torch.rand((Bi, 3, 720, 1280), generator=g).to(dev)
    with torch.no_grad():
        for _ in range(3):
            G(x)
        barrier()
        ev0.record()
        for _ in range(10):
            G(x)
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1)) / 10
        L.profile_enable(True)
        for _ in range(2):
            G(x)
        torch.cuda.synchronize()
        L.profile_enable(False)
    fam = L.profile_collect()
    d = fam["conv_tcgen05"]
    flops_all = sum(v["flops"] for v in fam.values()) / 2
    out["configs[4]_inference_1280x720_b4"] = {
        "images_per_s": world * Bi * 1e3 / ms, "ms_per_forward": ms, "per_gpu_batch": Bi, "n_gpus": world, "scaling": "replicas (weak)",
        "steps": 10, "warmup": 3, "algorithmic_tflops_whole_forward": flops_all / (ms / 1e3) / 1e12,
        "roofline": {"bound": "tensor", "kernel": "conv_tcgen05", "achieved": d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else None,
                     "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": (d["flops"] / (d["ms"] / 1e3) / 1e12 / pk["tf_sustained"]) if d["ms"] > 0 else None,
                     "share_of_profiled_kernel_time": d["ms"] / max(1e-9, sum(v["ms"] for v in fam.values()))}}
    del x
    torch.cuda.empty_cache()
    return out


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
    # per-launch durations need launches that do not overlap: the side-stream weight gradients (engine.ASYNC_WGRAD_MAX_PIXELS) are
    # switched off for this pass (the events between launches already serialise programmatic dependent launches)
    from fdgan_b200 import engine as _engine
    async_px, _engine.ASYNC_WGRAD_MAX_PIXELS = _engine.ASYNC_WGRAD_MAX_PIXELS, 0
    from fdgan_b200 import train as _train
    ovl, _train.OVERLAP_CLEAN_BRANCH = _train.OVERLAP_CLEAN_BRANCH, False      # likewise the clean-image branch on its auxiliary stream
    L.profile_enable(True)
    prof_steps = min(args.steps, 3)
    for i in range(prof_steps):
        h, c = resident[i % len(resident)]
        tr.step(h, c, sync_losses=False)
    torch.cuda.synchronize()
    L.profile_enable(False)
    _engine.ASYNC_WGRAD_MAX_PIXELS = async_px
    _train.OVERLAP_CLEAN_BRANCH = ovl
    fam = L.profile_collect()
    tot_ms = sum(v["ms"] for v in fam.values())
    dom = max(fam, key=lambda k: fam[k]["ms"])
    d = fam[dom]
    if d["ms"] > 0:
        prof, prof_src = ncu_family_profile(dom)
        if dom in ("conv_simt_f32", "conv_tcgen05", "wgrad"):
            achieved = d["flops"] / (d["ms"] / 1e3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tf_sustained"], "traffic": prof["dram_bytes_per_launch"] if prof else None, "kernel": dom,
                    "traffic_source": prof_src,
                    "algorithmic_bytes_per_launch": d["bytes"] / max(1, d["launches"]),
                    "traffic_over_algorithmic": (prof["dram_bytes_per_launch"] / (d["bytes"] / max(1, d["launches"]))) if prof and d["bytes"] > 0 else None,
                    "tensor_pipe_active_pct_ncu": prof["tensor_pipe_active_pct"] if prof else None,
                    "note": "fp32 operands run as bf16 hi/lo splits: 3 algorithmic bf16 passes per MAC (issued as 2 MMAs of width 2N and N), "
                            "so the algorithmic ceiling is 1/3 of the bf16 peak; achieved counts each MAC once",
                    "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                    "share_of_profiled_kernel_time": d["ms"] / tot_ms, "launches_per_step": d["launches"] / prof_steps,
                    "avg_launch_ms": d["ms"] / max(1, d["launches"]),
                    "flops_per_launch": d["flops"] / max(1, d["launches"])}
        else:
            achieved = d["bytes"] / (d["ms"] / 1e3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                    "traffic": prof["dram_bytes_per_launch"] if prof else None, "traffic_source": prof_src, "kernel": dom, "peak_source": pk["src"],
                    "share_of_profiled_kernel_time": d["ms"] / tot_ms}
    families = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                    "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 else 0.0,
                    "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 else 0.0,
                    "algorithmic_gb_per_step": v["bytes"] / prof_steps / 1e9} for k, v in fam.items() if v["launches"]}

    # ---- data-parallel self-check (N > 1): replicas identical, all-reduced gradient == single-process gradient of the same shards
    dp = None
    if world > 1:
        dp = dp_check(tr, world, rank, dev, size)

    # ---- secondary configs (outside the timed region): configs[1] B=1 step, configs[4] 1280x720 inference
    secondary = None
    if not args.no_secondary:
        secondary = secondary_configs(tr, G, world, rank, dev, size, pk, max_over_ranks, barrier)

    # ---- CPU baseline and the reference's torch / cuDNN path on this GPU (rank 0, N = 1 only)
    cpu = None
    gbase = None
    if world == 1 and not args.no_gpu_baseline:
        del resident
        torch.cuda.empty_cache()
        try:
            gbase = gpu_baseline(B, size, dev)
            gbase["fdgan_b200_over_strict_fp32"] = value / gbase["strict_fp32"]["value"]
            gbase["fdgan_b200_over_tf32"] = value / gbase["tf32"]["value"]
        except Exception as ex:      # a reported baseline must not take the bench line down
            gbase = {"unavailable": "%s: %s" % (type(ex).__name__, ex)}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        ips, s_per, cores, kind = cpu_step_images_per_s(2, size, 2, 1)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample_batch": 2,
               "sample": "2-image batches, 2 timed steps (+1 warm-up) of the full step, %s, %.1f s/step"
                         % ("reference modules from baseline/_ref via oracle/ref_step.py" if kind == "reference" else "oracle/fdgan_oracle.py:train_step", s_per)}

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
            "gpu_baseline": gbase,
            "secondary": secondary,
            "dp_check": dp,
            "losses_last_step": last,
            "notes": ["Blur / Laplacian (SURVEY 8 a8/a9) survive in the reference as bytecode only: their oracle is restated from the disassembly "
                      "and cross-checked against scipy -- parity unpinned for those two rows",
                      "loss terms / weights of the step are the SURVEY 3.3 reconstruction (the reference ships no train.py); the arithmetic of the "
                      "step is pinned to the reference's own modules + torch.optim.Adam (tests/golden/train_step_*.npz)"],
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
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-on-torch/cuDNN leg (N = 1 only)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1] / configs[4] legs")
    ap.add_argument("--ref-sample", type=int, default=4, help="--impl reference: images per timed CPU step (bounded sample)")
    ap.add_argument("--quick", action="store_true", help="value leg only (used under ncu): no e2e / profiled / CPU passes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
