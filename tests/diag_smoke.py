import sys; sys.path.insert(0, "/root/repo")
import torch, fdgan_b200
from fdgan_b200 import ops, engine
from oracle import fdgan_oracle as O
def run(umma, fused, k1, dtype=torch.float32):
    ops.USE_UMMA = umma; engine.FUSED_BN1_BWD = fused; ops.USE_K1 = k1
    sd = O.make_fdgan_state(0)
    net = fdgan_b200.FDGAN(); net.load_state_dict(sd); net = net.cuda().train()
    g = torch.Generator().manual_seed(5)
    x = torch.rand((1, 3, 32, 32), generator=g)
    sdd = {k: v.to(dtype) for k, v in sd.items()}
    for k in O.fdgan_used_param_names(): sdd[k].requires_grad_(True)
    yo = O.fdgan_forward(sdd, x.to(dtype), True, True); yo.square().mean().backward()
    y = net(x.cuda()); y.square().mean().backward(); torch.cuda.synchronize()
    err = float((y.detach().cpu().to(dtype) - yo.detach()).abs().max())
    out = []
    for k in ("dense_block1.denselayer1.conv1.weight", "dense_block3.denselayer24.conv2.weight", "conv_refin1.weight", "dense_block2.denselayer5.norm1.weight"):
        ga = dict(net.named_parameters())[k].grad.cpu().to(dtype); gb = sdd[k].grad
        out.append("%s max-rel %.2e relL2 %.2e" % (k.split(".")[-3] if k.count(".")>1 else k, float((ga-gb).abs().max()/gb.abs().max()), float((ga-gb).norm()/gb.norm())))
    print("umma=%d fused=%d k1=%d oracle=%s: fwd err %.2e | " % (umma, fused, k1, str(dtype)[-7:], err) + " | ".join(out), flush=True)
for dt in (torch.float32, torch.float64):
    run(False, False, False, dt); run(False, True, False, dt); run(True, False, False, dt); run(True, True, False, dt); run(True, True, True, dt)
