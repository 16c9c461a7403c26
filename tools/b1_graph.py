"""Batch-1 step replayed from a CUDA graph: ms per step (configs[1])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

def main():
    dev = torch.device("cuda:0")
    import fdgan_b200
    from fdgan_b200.train import GANTrainer
    torch.manual_seed(0)
    G, D, V = fdgan_b200.FDGAN(), fdgan_b200.D(9, 36), fdgan_b200.Vgg16()
    G, D, V = G.to(dev).train(), D.to(dev).train(), V.to(dev)
    for p in V.parameters():
        p.requires_grad_(False)
    tr = GANTrainer(G, D, V)
    h = torch.rand(1, 3, 256, 256, device=dev); c = torch.rand(1, 3, 256, 256, device=dev)
    for _ in range(3):
        tr.step_graphed(h, c, sync_losses=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        tr.step_graphed(h, c, sync_losses=False)
    e1.record(); torch.cuda.synchronize()
    print("batch-1 graphed step: %.3f ms (FDG_OVERLAP=%s)" % (e0.elapsed_time(e1) / 30, os.environ.get("FDG_OVERLAP", "1")))

main()
