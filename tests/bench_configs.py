"""Secondary BASELINE.json configs (not the bench.py headline): config[1] G+D adversarial step at B=1 256x256, and
config[4] 1280x720 inference at batch 4 (train-mode BatchNorm, no_grad).  python tests/bench_configs.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fdgan_b200
from fdgan_b200.train import GANTrainer


def ev_time(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
G, D, V = fdgan_b200.FDGAN().cuda().train(), fdgan_b200.D(9, 36).cuda().train(), fdgan_b200.Vgg16().cuda()
# config[1]: adversarial step, batch 1 (perceptual weight 0 => G + D only)
tr = GANTrainer(G, D, V, weights=dict(perc=0.0))
hz, cl = torch.rand(1, 3, 256, 256, device="cuda"), torch.rand(1, 3, 256, 256, device="cuda")
ms = ev_time(lambda: tr.step(hz, cl, sync_losses=False), 10)
print("config[1] G+D adversarial step, B=1 256x256: %.2f ms/step -> %.1f images/s" % (ms, 1e3 / ms), flush=True)
ms = ev_time(lambda: tr.step_graphed(hz, cl, sync_losses=False), 10)
print("           replayed from a CUDA graph (GANTrainer.step_graphed): %.2f ms/step -> %.1f images/s" % (ms, 1e3 / ms), flush=True)
tr2 = GANTrainer(G, D, V)
ms = ev_time(lambda: tr2.step(hz, cl, sync_losses=False), 10)
print("           with the VGG16 perceptual term, B=1: %.2f ms/step -> %.1f images/s" % (ms, 1e3 / ms), flush=True)
# config[4]: 1280x720 inference, batch 4
x = torch.rand(4, 3, 720, 1280, device="cuda")
with torch.no_grad():
    ms = ev_time(lambda: G(x), 3, warm=1)
print("config[4] FDGAN forward 1280x720 batch 4 (train-mode BN, no_grad): %.1f ms -> %.2f images/s, %.1f TFLOP/s algorithmic; peak memory %.1f GB"
      % (ms, 4e3 / ms, 4 * 1911.9e9 / ms / 1e9, torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
