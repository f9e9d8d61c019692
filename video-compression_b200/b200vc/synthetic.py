"""Synthetic workloads (no datasets or checkpoints exist offline): the seeded video generator of
SURVEY.md 8d config 2 and a deterministic re-scaling of random-init weights.

Why re-scale: with torch's default initialisation the analysis transforms shrink activations ~5x, so every
latent quantises to 0, every scale sits on the 0.11 clamp and SPyNet's flows are ~0.1 px -- the hot-path kernels
would only ever see one branch.  ``calibrate_`` multiplies a handful of final-layer weights by fixed constants
and perturbs GDN / factorised-prior parameters (seeded, CPU generator) so that flows reach a few pixels,
latents span several integer bins and scales cover the 64-entry table.  It is pure tensor arithmetic on the
state dict (no forward pass), so it applies identically to this package's ``lhbdc.Model`` and to any model
with the reference's key layout.
"""
import contextlib
import math

import torch
import torch.nn.functional as F


@contextlib.contextmanager
def _one_thread():
    """CPU ``F.interpolate`` rounds differently for different intra-op thread counts (vector-path boundaries move), and
    torchrun sets OMP_NUM_THREADS=1 per rank: build the canvas single-threaded so that a 1-process and an N-process
    run see bit-identical frames."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        yield
    finally:
        torch.set_num_threads(n)


@torch.no_grad()
def calibrate_(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    randn = lambda *s: torch.randn(*s, generator=g)
    rand = lambda *s: torch.rand(*s, generator=g)

    def put(p, value):
        p.copy_(value.to(device=p.device, dtype=p.dtype))

    for basic in model.FlowNet.netBasic:
        last = basic.netBasic[8]
        last.weight.mul_(3.0)
        last.bias.mul_(3.0)
    model.masknet.conv4.weight.mul_(60.0)
    for comp in (model.mv_compressor, model.residual_compressor):
        comp.g_a[6].weight.mul_(40.0)
        comp.g_a[6].bias.mul_(40.0)
        comp.h_a[8].weight.mul_(80.0)
        comp.h_s[8].weight.mul_(40.0)
        N = comp.h_s[8].weight.shape[0] // 2
        put(comp.h_s[8].bias[:N], torch.exp(math.log(0.05) + rand(N) * (math.log(20.0) - math.log(0.05))))
        for mod in comp.modules():
            if type(mod).__name__ == "GDN":
                C = mod.beta.numel()
                ped = mod.beta_reparam.pedestal.cpu()
                beta = 1.0 + 0.1 * randn(C).abs()
                gamma = 0.1 * torch.eye(C) + 0.01 * randn(C, C).abs()
                put(mod.beta, torch.sqrt(torch.max(beta + ped, ped)))
                put(mod.gamma, torch.sqrt(torch.max(gamma + ped, ped)))
        eb = comp.entropy_bottleneck
        for i in range(5):
            m = getattr(eb, f"_matrix{i}")
            put(m, m.cpu() + 0.2 * randn(*m.shape))
            if i < 4:
                f = getattr(eb, f"_factor{i}")
                put(f, 0.3 * randn(*f.shape))
        q = eb.quantiles.cpu().clone()
        q[:, 0, 1] = 0.5 * randn(q.shape[0])
        put(eb.quantiles, q)
    return model


@torch.no_grad()
def calibrate_flex_(model, seed=0):
    """Same idea for the Flex-Rate ``BidirFlowRef`` tree: flows of a few pixels, latents over several bins, gain
    matrices 1 + 0.1*randn (so level interpolation is exercised: SURVEY 8d config 3); un-zeroes the flow
    compressor's last layer (the reference zero-initialises it, which would make the refinement a no-op)."""
    g = torch.Generator().manual_seed(seed)
    randn = lambda *s: torch.randn(*s, generator=g)
    rand = lambda *s: torch.rand(*s, generator=g)

    def put(p, value):
        p.copy_(value.to(device=p.device, dtype=p.dtype))

    model.flow_predictor.last.weight.mul_(30.0)
    model.Mask.last.weight.mul_(30.0)
    for comp in (model.flow_compressor, model.residual_compressor):
        comp.g_a[6].weight.mul_(40.0)
        comp.g_a[6].bias.mul_(40.0)
        comp.h_a[8].weight.mul_(80.0)
        comp.h_s[8].weight.mul_(40.0)
        N = comp.h_s[8].weight.shape[0] // 2
        put(comp.h_s[8].bias[:N], torch.exp(math.log(0.05) + rand(N) * (math.log(20.0) - math.log(0.05))))
        for unit in (comp.gain_unit, comp.inv_gain_unit, comp.hyper_gain_unit, comp.hyper_inv_gain_unit):
            put(unit.gain_matrix, 1.0 + 0.1 * randn(*unit.gain_matrix.shape))
        for mod in comp.modules():
            if type(mod).__name__ == "GDN":
                C = mod.beta.numel()
                ped = mod.beta_reparam.pedestal.cpu()
                put(mod.beta, torch.sqrt(torch.max(1.0 + 0.1 * randn(C).abs() + ped, ped)))
                put(mod.gamma, torch.sqrt(torch.max(0.1 * torch.eye(C) + 0.01 * randn(C, C).abs() + ped, ped)))
        eb = comp.entropy_bottleneck
        for i in range(4):
            f = getattr(eb, f"_factor{i}")
            put(f, 0.3 * randn(*f.shape))
    last = model.flow_compressor.g_s[-1][0]
    put(last.weight, 0.02 * randn(*last.weight.shape))
    return model


def make_sequence(T, H=1080, W=1920, seed=1234, device="cpu", max_motion=8, noise=0.01):
    """[T,3,H,W] fp32 in [0,1]: a smooth random field (bicubic-upsampled uniform noise) seen through a window
    that drifts by up to ``max_motion`` px per frame, plus ``noise`` white noise.
    Generated on the CPU generator (same frames on every device), then moved."""
    g = torch.Generator().manual_seed(seed)
    margin = max_motion * 2 + 8
    ch, cw = (H + 2 * margin + 31) // 32, (W + 2 * margin + 31) // 32
    with _one_thread():
        coarse = torch.rand(1, 3, ch + 1, cw + 1, generator=g)
        canvas = F.interpolate(coarse, size=((ch + 1) * 32, (cw + 1) * 32), mode="bicubic", align_corners=False)
        fine = torch.rand(1, 3, (ch + 1) * 4, (cw + 1) * 4, generator=g)
        canvas = (0.8 * canvas + 0.2 * F.interpolate(fine, size=canvas.shape[-2:], mode="bilinear",
                                                     align_corners=False)).clamp_(0, 1)[0]
    frames = torch.empty(T, 3, H, W)
    for t in range(T):
        dx = margin + int(round(max_motion * math.sin(0.37 * t)))
        dy = margin + int(round(0.5 * max_motion * math.cos(0.23 * t)))
        win = canvas[:, dy:dy + H, dx:dx + W]
        frames[t] = (win + noise * torch.randn(win.shape, generator=g)).clamp_(0, 1)
    return frames.to(device)


def make_frames(indices, H=1080, W=1920, seed=1234, device="cpu", max_motion=8, noise=0.01):
    """Random-access variant of ``make_sequence`` for GOP-sharded runs: frame ``t`` depends on ``(seed, t)`` only
    (its noise comes from its own generator), so every rank can build exactly the frames of its shard and any two
    ranks agree bit for bit on a frame they both hold (the shared anchors).  The drifting canvas is the same
    construction as ``make_sequence``'s."""
    g = torch.Generator().manual_seed(seed)
    margin = max_motion * 2 + 8
    ch, cw = (H + 2 * margin + 31) // 32, (W + 2 * margin + 31) // 32
    with _one_thread():
        coarse = torch.rand(1, 3, ch + 1, cw + 1, generator=g)
        canvas = F.interpolate(coarse, size=((ch + 1) * 32, (cw + 1) * 32), mode="bicubic", align_corners=False)
        fine = torch.rand(1, 3, (ch + 1) * 4, (cw + 1) * 4, generator=g)
        canvas = (0.8 * canvas + 0.2 * F.interpolate(fine, size=canvas.shape[-2:], mode="bilinear",
                                                     align_corners=False)).clamp_(0, 1)[0]
    indices = list(indices)
    frames = torch.empty(len(indices), 3, H, W)
    for k, t in enumerate(indices):
        gt = torch.Generator().manual_seed(seed * 1000003 + 7919 * int(t) + 1)
        dx = margin + int(round(max_motion * math.sin(0.37 * t)))
        dy = margin + int(round(0.5 * max_motion * math.cos(0.23 * t)))
        win = canvas[:, dy:dy + H, dx:dx + W]
        frames[k] = (win + noise * torch.randn(win.shape, generator=gt)).clamp_(0, 1)
    return frames.to(device)


@torch.no_grad()
def calibrate_flowguided_(model, seed=0):
    """Same idea for the ICIP2024 ``FlowGuidedB`` tree (reference attribute names: ICIP2024/src/model/m.py:31-49):
    flows of a few pixels at the pooled resolution, latents over several bins through the gain tables (distinct per
    rate level so that level interpolation matters), predicted scales over the table, perturbed factorised priors."""
    g = torch.Generator().manual_seed(seed)
    randn = lambda *s: torch.randn(*s, generator=g)
    rand = lambda *s: torch.rand(*s, generator=g)

    def put(p, value):
        p.copy_(value.to(device=p.device, dtype=p.dtype))

    last = model.flow_estimator.up3[-1][0]
    last.weight.mul_(12.0)
    last.bias.mul_(12.0)
    for comp, base in ((model.offset_compressor, 6.0), (model.residual_compressor, 10.0)):
        L, M = comp.Gain.shape
        lvl = torch.linspace(0.6, 1.6, L).view(L, 1)
        gain = base * lvl * (1.0 + 0.1 * randn(L, M))
        put(comp.Gain, gain)
        put(comp.InverseGain, (1.0 + 0.05 * randn(L, M)) / gain)
        hgain = 3.0 * lvl * (1.0 + 0.1 * randn(L, comp.HyperGain.shape[1]))
        put(comp.HyperGain, hgain)
        put(comp.InverseHyperGain, (1.0 + 0.05 * randn(L, comp.HyperGain.shape[1])) / hgain)
        for ep in comp.entropy_parameters:
            c = ep[-1].bias.shape[0] // 2
            put(ep[-1].bias[:c], torch.exp(math.log(0.05) + rand(c) * (math.log(20.0) - math.log(0.05))))
        eb = comp.entropy_bottleneck
        for i in range(4):
            f = getattr(eb, f"_factor{i}")
            put(f, 0.3 * randn(*f.shape))
        q = eb.quantiles.cpu().clone()
        q[:, 0, 1] = 0.5 * randn(q.shape[0])
        put(eb.quantiles, q)
    return model
