"""GPU path against golden outputs of the reference itself (tests/golden, made by oracle/make_golden.py) -- no oracle in between."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_experts_vs_reference_golden_64():
    from isr2_b200 import model as M
    g = torch.load(os.path.join(GOLD, "experts_64x64_fp16.pt"))
    m = M.FreqFusionB200("cuda:0", init_seed=0, verbose=False)
    ex = m.expert_outputs_nchw(g["x"].cuda())
    for k in ("hat", "dat", "nafnet"):
        err = (ex[k].cpu() - g[k].float()).abs().max().item()
        assert err < 2e-2, f"{k}: {err}"


def test_head_vs_reference_golden_64():
    from isr2_b200 import head, weights
    g = torch.load(os.path.join(GOLD, "head_64.pt"))
    lr = g["lr"]
    gen = torch.Generator().manual_seed(g["expert_seed"])
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    ex = [(up + s * torch.randn(1, 3, 256, 256, generator=gen)).clamp(0, 1) for s in (0.01, 0.02, 0.03)]
    stack = torch.zeros(256 * 256, 12)
    stack[:, :9] = torch.cat(ex, 1).permute(0, 2, 3, 1).reshape(-1, 9)
    r = head.HeadRunner(weights.make_state_dict("fusion", 0), torch.device("cuda:0"))
    inter = {}
    out = r.forward(lr.cuda(), stack.cuda(), intermediates=inter).cpu()
    nh = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
    assert (inter["bands_raw"].cpu() - nh(g["raw_bands"])).abs().max() < 1e-4       # fp32 DCT/DWT/DFT kernels
    assert (inter["band_features"].cpu() - nh(g["band_features"])).abs().max() < 2e-2
    assert (out - g["out"]).abs().max() < 2e-2


def test_forward_with_precomputed_vs_reference_golden_64():
    """BASELINE.json configs[0]: fusion head alone on synthetic expert outputs, public API."""
    from isr2_b200 import model as M
    g = torch.load(os.path.join(GOLD, "head_64.pt"))
    lr = g["lr"]
    gen = torch.Generator().manual_seed(g["expert_seed"])
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    ex = {k: (up + s * torch.randn(1, 3, 256, 256, generator=gen)).clamp(0, 1).cuda() for k, s in (("hat", 0.01), ("grl", 0.02), ("nafnet", 0.03))}
    m = M.FreqFusionB200("cuda:0", init_seed=0, verbose=False)
    out = m.forward_with_precomputed(lr.cuda(), ex).cpu()
    assert (out - g["out"]).abs().max() < 2e-2
    with pytest.raises(KeyError):
        m.forward_with_precomputed(lr.cuda(), {"hat": ex["hat"]})


def test_cached_expert_files_through_the_head(tmp_path):
    """The reference's cached-expert workflow (src/data/cached_dataset.py + forward_with_precomputed): the golden head input
    written in the on-disk format ({stem}_drct_part.pt with keys drct / grl + {stem}_rest_part.pt), read back by
    isr2_b200.cached and run through the fusion head; SR must match the reference's golden output, and the on-GPU PSNR / SSIM
    against the cached HR patch must match the oracle metrics of the same SR."""
    from isr2_b200 import cached, model as M
    from oracle import metrics
    g = torch.load(os.path.join(GOLD, "head_64.pt"))
    lr = g["lr"]
    gen = torch.Generator().manual_seed(g["expert_seed"])
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    ex = {k: (up + s * torch.randn(1, 3, 256, 256, generator=gen)).clamp(0, 1) for k, s in (("hat", 0.01), ("grl", 0.02), ("nafnet", 0.03))}
    hr = up.clamp(0, 1)[0]
    for stem in ("a_p0", "a_p1"):
        torch.save({"outputs": {"drct": ex["hat"]}, "lr": lr[0], "hr": hr, "filename": stem}, tmp_path / f"{stem}_drct_part.pt")
        torch.save({"outputs": {"grl": ex["grl"], "nafnet": ex["nafnet"]}, "filename": stem}, tmp_path / f"{stem}_rest_part.pt")
    store = cached.CachedExpertStore(str(tmp_path))
    assert len(store) == 2
    m = M.FreqFusionB200("cuda:0", init_seed=0, verbose=False)
    recs = cached.run_cached(m, store, batch_size=2)
    assert [r["filename"] for r in recs] == ["a_p0", "a_p1"]
    for r in recs:
        sr = r["sr"].cpu()[None]
        assert (sr - g["out"]).abs().max() < 2e-2
        assert abs(r["psnr_y"] - metrics.psnr_y(sr, hr[None])[0]) < 1e-3
        ref_ssim = metrics.ssim_y(sr, hr[None])[0]
        print(f"cached: psnr_y {r['psnr_y']:.4f} dB, ssim_y {r['ssim_y']:.7f} (oracle {ref_ssim:.7f})")
        assert abs(r["ssim_y"] - ref_ssim) < 2e-5      # smooth image: small variances in the denominator amplify fp32 summation-order noise


def _collab_inputs(g):
    from oracle import collab
    lr = g["lr"]
    gen = torch.Generator().manual_seed(g["expert_seed"])
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    ex = {k: (up + s * torch.randn(1, 3, 256, 256, generator=gen)).clamp(0, 1) for k, s in (("hat", 0.01), ("dat", 0.02), ("nafnet", 0.03))}
    return lr, ex, collab.synth_features(1, 64, 64, g["feature_seed"])


def test_forward_with_precomputed_and_expert_features_vs_reference_golden():
    """The reference's cached-mode path WITH expert features: forward_with_precomputed(lr, outputs, features) runs
    EnhancedCollaborativeWithLKA (align 1x1 convs -> 3-token cross-expert attention -> FFN -> shared LKA block -> per-expert
    modulation heads) before the fusion head (enhanced_fusion.py:466-496, large_kernel_attention.py:251-419).  Golden output of
    the reference itself: tests/golden/head_collab_64.pt (oracle/make_golden_collab.py)."""
    from isr2_b200 import model as M
    g = torch.load(os.path.join(GOLD, "head_collab_64.pt"))
    lr, ex, feats = _collab_inputs(g)
    m = M.FreqFusionB200("cuda:0", init_seed=0, verbose=False)
    out = m.forward_with_precomputed(lr.cuda(), {k: v.cuda() for k, v in ex.items()}, {k: v.cuda() for k, v in feats.items()}).cpu()
    stack = m._stack(1, 64, 64).cpu().view(1, 256, 256, 12)[..., :9].permute(0, 3, 1, 2)
    e_enh = (stack - g["enhanced"].float()).abs().max().item()
    e_out = (out - g["out"]).abs().max().item()
    print(f"collaborative: enhanced expert outputs {e_enh:.2e}, final output {e_out:.2e}")
    assert e_enh < 2e-3 and e_out < 2e-2
    # an expert without features contributes zeros to the cross-expert attention (large_kernel_attention.py:374-377): against the oracle
    from oracle import collab
    only_hat = {"hat": feats["hat"]}
    with torch.no_grad():
        ref = torch.cat(collab.collaborative(m.state["fusion"], only_hat, [ex["hat"], ex["dat"], ex["nafnet"]]), 1)
    m.forward_with_precomputed(lr.cuda(), {k: v.cuda() for k, v in ex.items()}, {"hat": feats["hat"].cuda()})
    stack = m._stack(1, 64, 64).cpu().view(1, 256, 256, 12)[..., :9].permute(0, 3, 1, 2)
    assert (stack - ref).abs().max().item() < 2e-3
    with pytest.raises(Exception):      # the smallest feature map must have the LR size
        m.forward_with_precomputed(lr.cuda(), {k: v.cuda() for k, v in ex.items()}, {"hat": feats["hat"][..., :32, :32].cuda()})


def test_collaborative_feature_alignment_paths_vs_reference_golden():
    """The alignment step of EnhancedCollaborativeWithLKA.forward (large_kernel_attention.py:337-378) on features that take every branch
    of it -- 200 channels for hat (truncated), 150 for dat (zero padded), nafnet at twice the spatial size (resized to the smallest) --
    against the reference itself: tests/golden/head_collab_mixed_32.pt (oracle/make_golden_collab.py)."""
    from isr2_b200 import model as M
    from oracle import collab
    g = torch.load(os.path.join(GOLD, "head_collab_mixed_32.pt"))
    lr = g["lr"]
    gen = torch.Generator().manual_seed(g["expert_seed"])
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    ex = {k: (up + s * torch.randn(1, 3, 128, 128, generator=gen)).clamp(0, 1) for k, s in (("hat", 0.01), ("dat", 0.02), ("nafnet", 0.03))}
    feats = collab.synth_features_mixed(1, 32, 32, g["feature_seed"])
    m = M.FreqFusionB200("cuda:0", init_seed=0, verbose=False)
    out = m.forward_with_precomputed(lr.cuda(), {k: v.cuda() for k, v in ex.items()}, {k: v.cuda() for k, v in feats.items()}).cpu()
    stack = m._stack(1, 32, 32).cpu().view(1, 128, 128, 12)[..., :9].permute(0, 3, 1, 2)
    e_enh = (stack - g["enhanced"].float()).abs().max().item()
    e_out = (out - g["out"]).abs().max().item()
    print(f"collaborative, mixed features: enhanced expert outputs {e_enh:.2e}, final output {e_out:.2e}")
    assert e_enh < 2e-3 and e_out < 2e-2


def test_collaborative_modulation_vs_oracle_with_strong_heads():
    """Same branch against the oracle with the modulation heads' last layer scaled up (the benign factory leaves every modulation
    within 1 % of 1.0, too little to tell a wrong attention / LKA / pooling from a right one): the per-expert, per-channel
    modulation values themselves must agree."""
    from isr2_b200 import model as M
    from oracle import collab
    g = torch.load(os.path.join(GOLD, "head_collab_64.pt"))
    lr, ex, feats = _collab_inputs(g)
    m = M.FreqFusionB200("cuda:0", init_seed=0, verbose=False)
    for i in range(3):
        m.state["fusion"][f"collaborative.modulation.{i}.3.weight"] = m.state["fusion"][f"collaborative.modulation.{i}.3.weight"] * 6
        m.state["fusion"][f"collaborative.modulation.{i}.0.weight"] = m.state["fusion"][f"collaborative.modulation.{i}.0.weight"] * 3
    m._runners = None
    sd = m.state["fusion"]
    exl = [ex["hat"], ex["dat"], ex["nafnet"]]
    with torch.no_grad():
        _, ref_mod = collab.collaborative(sd, feats, exl, return_mod=True)      # [1, 3 experts, 3 channels]
        ref = collab.head_forward_with_features(sd, lr, exl, feats)
    inter = {}
    out = m.forward_with_precomputed(lr.cuda(), {k: v.cuda() for k, v in ex.items()}, {k: v.cuda() for k, v in feats.items()}, intermediates=inter).cpu()
    got_mod = inter["modulation"].cpu()
    print("modulation (oracle):", [f"{v:.3f}" for v in ref_mod.flatten().tolist()])
    print("modulation (ours):  ", [f"{v:.3f}" for v in got_mod.flatten().tolist()])
    assert (ref_mod - 0.5).abs().max() > 0.05                      # the test has teeth
    assert (got_mod - ref_mod).abs().max().item() < 1e-2
    assert (out - ref).abs().max().item() < 2e-2
