"""GPU parity: K-SPYPYR / K-SPYLEVEL through the C-ABI vs the torch chain of
LHBDC/model/flow.py:78-101 restated in oracle/lhbdc.py (Preprocess, avg_pool2d pyramid, x2 align_corners=True flow
upsample * 2, replicate pad, backwarp, concat), run on the same device.  Bar: bit-exact (these are the same fp32
operations in the same order); the assertion allows 1e-6 absolute so that a one-ulp difference would be reported
with its size instead of a bare mismatch."""
import pytest
import torch
import torch.nn.functional as F

from oracle import lhbdc as o_lhbdc
from oracle import warp as o_warp

pytestmark = pytest.mark.gpu


def _torch_pyramid(x):
    lv = [o_lhbdc._Preprocess()(x)]
    for _ in range(5):
        if lv[0].shape[2] > 32 or lv[0].shape[3] > 32:
            lv.insert(0, F.avg_pool2d(lv[0], kernel_size=2, stride=2, count_include_pad=False))
    return lv


def _torch_level(a, b, flow):
    up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
    if up.shape[2] != a.shape[2]:
        up = F.pad(up, [0, 0, 0, 1], mode="replicate")
    if up.shape[3] != a.shape[3]:
        up = F.pad(up, [0, 1, 0, 0], mode="replicate")
    return torch.cat([a, o_warp.backwarp_lhbdc(b, up), up], 1)


@pytest.mark.parametrize("shape", [(1, 192, 256), (2, 70, 131), (1, 24, 30), (1, 33, 20), (3, 64, 64), (1, 1088, 1920)])
def test_pyramid_is_bit_exact(shape):
    from b200vc import ops
    N, H, W = shape
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = torch.rand(N, 3, H, W, generator=g).cuda()
    want = _torch_pyramid(x)
    got = ops.spynet_pyramid(x)
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in want]
    for lvl, (a, b) in enumerate(zip(got, want)):
        exact = (a == b).float().mean().item()
        err = (a - b).abs().max().item()
        print(f"pyramid {shape} level {lvl} {tuple(a.shape)}: max|diff|={err:.3e} bit-exact={exact:.5f}")
        assert torch.equal(a, b), (lvl, err)


def test_pyramid_accepts_a_channel_slice_and_batch_stride():
    from b200vc import ops
    big = torch.rand(2, 7, 96, 80, device="cuda")
    x = big[:, 2:5]
    for a, b in zip(ops.spynet_pyramid(x), _torch_pyramid(x.contiguous())):
        assert torch.equal(a, b)


# >= 2 Mpx per launch with W % 4 == 0 takes the TMA-staged kernel (warp_tma.cu), everything else the gather kernel
@pytest.mark.parametrize("shape", [(1, 34, 60), (2, 68, 120), (1, 17, 31), (1, 35, 64), (2, 136, 240), (1, 544, 960),
                                   (1, 1088, 1920), (4, 544, 960), (3, 1087, 1924)])
@pytest.mark.parametrize("amp", [0.7, 6.0])
def test_level_is_bit_exact(shape, amp):
    from b200vc import ops
    N, H, W = shape
    g = torch.Generator().manual_seed(H * 7 + W + int(amp * 10))
    a = torch.randn(N, 3, H, W, generator=g).cuda()
    b = torch.randn(N, 3, H, W, generator=g).cuda()
    flow = (amp * torch.randn(N, 2, H // 2, W // 2, generator=g)).cuda()
    want = _torch_level(a, b, flow)
    got = ops.spynet_level(a, b, flow)
    assert got.shape == want.shape
    for name, sl in (("first", slice(0, 3)), ("warped", slice(3, 6)), ("up", slice(6, 8))):
        err = (got[:, sl] - want[:, sl]).abs().max().item()
        exact = (got[:, sl] == want[:, sl]).float().mean().item()
        print(f"level {shape} amp={amp} {name}: max|diff|={err:.3e} bit-exact={exact:.5f}")
        assert err <= 1e-6, (name, err)
    assert torch.equal(got[:, 0:3], want[:, 0:3]) and torch.equal(got[:, 6:8], want[:, 6:8])


def test_level_zero_flow_is_the_reference_start():
    from b200vc import ops
    a = torch.randn(2, 3, 34, 60, device="cuda")
    b = torch.randn(2, 3, 34, 60, device="cuda")
    want = _torch_level(a, b, a.new_zeros(2, 2, 17, 30))
    got = ops.spynet_level(a, b, None)
    assert torch.equal(got, want)


def test_level_rejects_a_flow_that_does_not_upsample_to_the_level():
    from b200vc import ops
    a = torch.randn(1, 3, 34, 60, device="cuda")
    with pytest.raises(RuntimeError):
        ops.spynet_level(a, a, torch.zeros(1, 2, 16, 30, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.spynet_level(a, a[:, :2], None)


@pytest.mark.parametrize("shape", [(1, 192, 256), (2, 128, 192), (1, 1088, 1920)])
def test_flownet_matches_oracle_network(shape, strict_fp32):
    """Whole SPyNet (fused glue + cuDNN convs) vs the oracle network: same weights, same device."""
    import b200vc
    N, H, W = shape
    torch.manual_seed(3)
    orc = o_lhbdc.Network().cuda().eval()
    prod = b200vc.lhbdc.Network().cuda().eval()
    prod.load_state_dict(orc.state_dict())
    g = torch.Generator().manual_seed(11)
    x1 = torch.rand(N, 3, H, W, generator=g).cuda()
    x2 = (x1 + 0.05 * torch.randn(N, 3, H, W, generator=g).cuda()).clamp(0, 1)
    with torch.no_grad():
        want = orc(x1, x2)
        got = prod(x1, x2)
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    print(f"FlowNet {shape}: max|diff|={err:.3e} (flow magnitude {scale:.3e}) bit-exact={(got == want).float().mean().item():.5f}")
    assert err <= 1e-5 * max(1.0, scale)


def test_level_gather_kernel_at_full_hd():
    """The gather kernel on the shape the staged kernel normally takes (B200VC_SPYNET_TMA=0, own process)."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, torch\n"
        "sys.path[:0] = [%r, %r, %r]\n"
        "import test_gpu_spynet as t\n"
        "from b200vc import ops\n"
        "g = torch.Generator().manual_seed(5)\n"
        "a = torch.randn(2, 3, 1088, 1920, generator=g).cuda(); b = torch.randn(2, 3, 1088, 1920, generator=g).cuda()\n"
        "f = (3 * torch.randn(2, 2, 544, 960, generator=g)).cuda()\n"
        "assert torch.equal(ops.spynet_level(a, b, f), t._torch_level(a, b, f))\n"
        "print('ok')\n"
    ) % (os.path.dirname(__file__), os.path.dirname(os.path.dirname(__file__)),
         os.path.join(os.path.dirname(os.path.dirname(__file__)), "video-compression_b200"))
    env = dict(os.environ, B200VC_SPYNET_TMA="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
