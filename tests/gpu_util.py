import math

import torch


def rel_err(got, want, floor=1e-12):
    return ((got.double() - want.double()).abs() / want.double().abs().clamp(min=floor)).max().item()


def warp_case(seed, N, C, H, W, amp=3.0, device="cuda"):
    """SURVEY 8d value distributions: image U[0,1]; flow N(0, amp^2) px with ~1 % of vectors out of frame."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(N, C, H, W, generator=g)
    flow = amp * torch.randn(N, 2, H, W, generator=g)
    far = torch.rand(N, 1, H, W, generator=g) < 0.01
    flow = torch.where(far, flow * 60.0, flow)
    return img.to(device), flow.to(device)


def gc_case(seed, N, C, H, W, device="cuda"):
    """mu ~ N(0,2), sigma ~ logU[0.05, 64] (hits the 0.11 clamp and every table bin), y = mu + sigma*N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    mu = 2.0 * torch.randn(N, C, H, W, generator=g)
    sigma = torch.exp(math.log(0.05) + torch.rand(N, C, H, W, generator=g) * (math.log(64.0) - math.log(0.05)))
    y = mu + sigma * torch.randn(N, C, H, W, generator=g)
    return y.to(device), sigma.to(device), mu.to(device)


def build_models(device="cuda", seed=0):
    """Oracle model (torch ops) and product model (b200vc kernels) with identical calibrated weights."""
    import b200vc
    from b200vc import synthetic
    from oracle import lhbdc as o_lhbdc
    torch.manual_seed(seed)
    orc = o_lhbdc.Model().eval()
    synthetic.calibrate_(orc, 0)
    prod = b200vc.Model().eval()
    prod.load_state_dict(orc.state_dict())
    for m in (orc, prod):
        m.mv_compressor.update(force=True)
        m.residual_compressor.update(force=True)
    return orc.to(device), prod.to(device)
