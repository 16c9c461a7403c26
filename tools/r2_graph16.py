import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdgan_b200, bench
from fdgan_b200.train import GANTrainer
torch.manual_seed(0)
G, D, V = fdgan_b200.FDGAN().cuda().train(), fdgan_b200.D(9, 36).cuda().train(), fdgan_b200.Vgg16().cuda()
for p in V.parameters(): p.requires_grad_(False)
tr = GANTrainer(G, D, V)
B = int(os.environ.get("B", "16"))
h, c = [t.cuda() for t in bench.make_batches(1, B, 256, 0, False)[0]]
def tm(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("B=%d eager   %.2f ms" % (B, tm(lambda: tr.step(h, c, sync_losses=False), 10)))
print("B=%d graphed %.2f ms" % (B, tm(lambda: tr.step_graphed(h, c, sync_losses=False), 10)))
print("peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
