"""Reported baseline, not a product path: the oracle's restatement of the reference arithmetic (plain torch.nn.functional,
oracle/fdgan_oracle.py:train_step) executed by eager PyTorch / cuDNN on the SAME GPU, i.e. what the reference's own
modules would run as on this box (`/root/reference` does not travel to the GPU box; SURVEY 8d "Reference-on-GPU
baseline").  BASELINE.json's target is >= 6x images/s over this path at batch 16, 256x256.

    python tests/bench_torch_eager.py [--batch 16] [--size 256] [--steps 5] [--device cuda]

Prints one line per precision mode: strict fp32 (TF32 off: the reference's arithmetic) and TF32 allowed (torch's cuDNN
default).  `cudnn.benchmark = True` as in demo.py:11."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import fdgan_oracle as O


def synth_pair(i, size):
    g = torch.Generator().manual_seed(1234 + i)     # J ~ U[0,1), I = J t + A (1 - t)
    j = torch.rand(3, size, size, generator=g)
    t = 0.3 + 0.6 * torch.rand((), generator=g)
    a = 0.7 + 0.3 * torch.rand((), generator=g)
    return j * t + a * (1 - t), j


def to_dev(sd, dev):
    return type(sd)((k, v.to(dev)) for k, v in sd.items())


def run(batch, size, steps, warmup, dev, tf32):
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    g_sd, d_sd, v_sd = (to_dev(s, dev) for s in (O.make_fdgan_state(0), O.make_d_state(9, 36, 1), O.make_vgg_state(2)))
    pairs = [synth_pair(i, size) for i in range(batch)]
    hazy, clean = torch.stack([p[0] for p in pairs]).to(dev), torch.stack([p[1] for p in pairs]).to(dev)
    sg, sd = {}, {}
    cuda = dev.type == "cuda"
    for _ in range(warmup):
        O.train_step(g_sd, d_sd, v_sd, hazy, clean, sg, sd)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(g_sd, d_sd, v_sd, hazy, clean, sg, sd)
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    else:
        ms = 1e3 * (time.perf_counter() - t0) / steps
    mem = torch.cuda.max_memory_allocated() / 2 ** 30 if cuda else 0.0
    print("torch eager %s, %s: batch %d %dx%d: %.1f ms/step -> %.1f images/s (peak memory %.1f GB)"
          % ("cuDNN" if cuda else "CPU", "TF32 allowed" if tf32 else "strict fp32", batch, size, size, ms, 1e3 * batch / ms, mem), flush=True)
    return ms


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    dev = torch.device(a.device)
    for tf32 in (False, True):
        run(a.batch, a.size, a.steps, a.warmup, dev, tf32)
