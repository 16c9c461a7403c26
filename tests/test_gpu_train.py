"""GPU parity of the training step (fdgan_b200.train.GANTrainer) against the oracle's reconstructed step."""
import pytest
import torch

from oracle import fdgan_oracle as O
from tests.util import grad_close, maxabs, seeded

pytestmark = pytest.mark.gpu


def _nets():
    import fdgan_b200
    G, D, V = fdgan_b200.FDGAN(), fdgan_b200.D(9, 36), fdgan_b200.Vgg16()
    G.load_state_dict(O.make_fdgan_state(0))
    D.load_state_dict(O.make_d_state(9, 36, 1))
    V.load_state_dict(O.make_vgg_state(2))
    return G.cuda().train(), D.cuda().train(), V.cuda()


@pytest.mark.parametrize("batch,hw", [(1, 32), (2, 48)])
def test_train_step_matches_oracle(batch, hw):
    from fdgan_b200.train import GANTrainer
    G, D, V = _nets()
    tr = GANTrainer(G, D, V)
    g_sd, d_sd, v_sd = O.make_fdgan_state(0), O.make_d_state(9, 36, 1), O.make_vgg_state(2)
    sg, sdd = {}, {}
    hazy, clean = seeded((batch, 3, hw, hw), 5), seeded((batch, 3, hw, hw), 6)
    for it in range(2):
        parts, gd, gg, fake_o = O.train_step(g_sd, d_sd, v_sd, hazy, clean, sg, sdd)
        fake = tr.step(hazy.cuda(), clean.cuda())
        # step 2 runs on parameters that went through Adam's first, sign-like update (lr * g / (|g| + eps)), which
        # amplifies rounding-level gradient differences into +-2*lr parameter differences
        assert maxabs(fake, fake_o) <= (2e-4 if it == 0 else 3e-2)
        for k in ("loss_d", "loss_g", "l1_weighted", "perc_weighted", "adv_weighted"):
            assert abs(tr.last[k] - parts[k]) <= (2e-3 if it == 0 else 2e-2) * max(1.0, abs(parts[k])), (it, k, tr.last[k], parts[k])
        if it == 0:
            for k, g in gd.items():
                grad_close(tr.sD.grad_views[k], g, "D " + k, rel_l2=6e-2, rel_max=0.5)
            worst = 0.0
            for k, g in gg.items():
                l2, _ = grad_close(tr.sG.grad_views[k], g, "G " + k, rel_l2=6e-2, rel_max=0.5)
                worst = max(worst, l2)
            print("worst G-gradient rel-L2", worst)
    # parameters moved by Adam exactly like the oracle's (first step: lr * sign-like update)
    moved = maxabs(dict(D.named_parameters())["main.layer5.conv.weight"], d_sd["main.layer5.conv.weight"])
    assert moved <= 1e-3


def test_train_step_with_ssim_term_matches_oracle():
    """The generator loss with the SSIM term: w_ssim * (1 - ssim(fake, clean)), pytorch_ssim semantics."""
    from fdgan_b200.train import GANTrainer
    G, D, V = _nets()
    wts = dict(ssim=0.4)
    tr = GANTrainer(G, D, V, weights=wts)
    g_sd, d_sd, v_sd = O.make_fdgan_state(0), O.make_d_state(9, 36, 1), O.make_vgg_state(2)
    hazy, clean = seeded((2, 3, 48, 48), 5), seeded((2, 3, 48, 48), 6)
    parts, gd, gg, fake_o = O.train_step(g_sd, d_sd, v_sd, hazy, clean, {}, {}, weights=wts)
    fake = tr.step(hazy.cuda(), clean.cuda())
    assert maxabs(fake, fake_o) <= 2e-4
    assert abs(tr.last["loss_g"] - parts["loss_g"]) <= 2e-3 * max(1.0, abs(parts["loss_g"])), (tr.last, parts)
    want_ssim = 0.4 * (1 - float(O.ssim(fake_o, clean)))
    assert abs(tr.last["ssim_weighted"] - want_ssim) <= 1e-4
    for k, g in gg.items():
        grad_close(tr.sG.grad_views[k], g, "G " + k, rel_l2=6e-2, rel_max=0.5)


def test_graphed_step_matches_eager_step():
    """GANTrainer.step_graphed (CUDA-graph replay, device-side Adam step counter) == GANTrainer.step over three iterations."""
    from fdgan_b200.train import GANTrainer
    hazy, clean = seeded((1, 3, 64, 64), 5).cuda(), seeded((1, 3, 64, 64), 6).cuda()
    hazy2, clean2 = seeded((1, 3, 64, 64), 7).cuda(), seeded((1, 3, 64, 64), 8).cuda()
    Ga, Da, Va = _nets()
    Gb, Db, Vb = _nets()
    ta, tb = GANTrainer(Ga, Da, Va), GANTrainer(Gb, Db, Vb)
    for it, (h, c) in enumerate(((hazy, clean), (hazy2, clean2), (hazy, clean2))):
        fa = ta.step(h, c).clone()
        fb = tb.step_graphed(h, c).clone()
        assert maxabs(fb, fa) <= (1e-5 if it == 0 else 2e-2), it      # later steps: split-K atomics order -> Adam sign-like amplification
        for k in ("loss_d", "loss_g"):
            assert abs(ta.last[k] - tb.last[k]) <= (1e-5 if it == 0 else 2e-2) * max(1.0, abs(ta.last[k])), (it, k)
    assert tb.sG.step == 3 and abs(float(tb.sG.dev_state[0]) - 3.0) < 1e-6
    pa, pb = dict(Da.named_parameters())["main.layer5.conv.weight"], dict(Db.named_parameters())["main.layer5.conv.weight"]
    assert maxabs(pb, pa) <= 2e-3


def test_step_is_the_same_with_and_without_stream_overlap(monkeypatch):
    """The two auxiliary-stream forks of GANTrainer.step (clean-image branch beside the generator forward, perceptual branch beside the
    discriminator step) only reorder independent work: first step identical up to the fp64-atomics order of the statistics."""
    from fdgan_b200 import train
    from fdgan_b200.train import GANTrainer
    hazy, clean = seeded((2, 3, 64, 64), 5).cuda(), seeded((2, 3, 64, 64), 6).cuda()
    res = []
    for on in (True, False):
        monkeypatch.setattr(train, "OVERLAP_CLEAN_BRANCH", on)
        monkeypatch.setattr(train, "OVERLAP_PERC_BRANCH", on)
        G, D, V = _nets()
        tr = GANTrainer(G, D, V)
        fake = tr.step(hazy, clean).clone()
        torch.cuda.synchronize()
        res.append((fake, dict(tr.last), tr.sG.grad.clone(), tr.sD.grad.clone()))
    (fa, la, ga, da), (fb, lb, gb, db) = res
    assert maxabs(fa, fb) <= 1e-6
    for k in ("loss_d", "l1_weighted", "perc_weighted", "adv_weighted", "loss_g"):
        assert abs(la[k] - lb[k]) <= 1e-6 * max(1.0, abs(la[k])), k
    assert float((ga - gb).norm() / gb.norm()) <= 1e-4 and float((da - db).norm() / db.norm()) <= 1e-4


def test_trainer_uses_flat_buffers():
    from fdgan_b200.train import GANTrainer
    G, D, V = _nets()
    tr = GANTrainer(G, D, V)
    for n, p in G._used_named_parameters():
        assert p.data_ptr() >= tr.sG.flat.data_ptr() and p.data_ptr() < tr.sG.flat.data_ptr() + 4 * tr.sG.n
    assert tr.sG.n >= 11803155 and tr.sD.n >= 790416
