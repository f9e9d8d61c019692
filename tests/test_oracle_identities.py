"""CPU: closed-form identities pinning the CompressAI restatement (parity is otherwise unpinned at that
boundary: compressai==1.1.8 is not installable offline and the reference holds no vectors for it)."""
import math

import torch

from oracle import cai


def test_gdn_at_init_is_closed_form():
    x = torch.randn(2, 8, 5, 7) * 3
    g, ig = cai.GDN(8), cai.GDN(8, inverse=True)
    torch.testing.assert_close(g(x), x / torch.sqrt(1 + 0.1 * x ** 2), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ig(x), x * torch.sqrt(1 + 0.1 * x ** 2), rtol=1e-5, atol=1e-6)
    keys = set(g.state_dict())
    assert keys == {"beta", "gamma", "beta_reparam.pedestal", "beta_reparam.lower_bound.bound",
                    "gamma_reparam.pedestal", "gamma_reparam.lower_bound.bound"}


def test_gaussian_conditional_matches_normal_cdf_difference():
    gc = cai.GaussianConditional(None).eval()
    y = torch.randn(4, 6, 9, 9) * 5
    mu = torch.randn_like(y)
    sigma = torch.exp(torch.empty_like(y).uniform_(math.log(0.05), math.log(64)))
    y_hat, lik = gc(y, sigma, means=mu)
    assert torch.equal(y_hat, torch.round(y - mu) + mu)
    s = sigma.clamp(min=0.11).double()
    n = torch.distributions.Normal(mu.double(), s)
    want = (n.cdf(y_hat.double() + 0.5) - n.cdf(y_hat.double() - 0.5)).clamp(min=1e-9)
    assert (lik.double() - want).abs().max().item() < 5e-7
    assert lik.min().item() >= float(torch.tensor(1e-9, dtype=torch.float32))


def test_round_is_half_to_even_and_symbols_are_ints():
    gc = cai.GaussianConditional(None)
    v = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5])
    assert gc.quantize(v, "symbols").tolist() == [0, 2, 2, 0, -2]
    assert gc.quantize(v, "symbols").dtype == torch.int32


def test_build_indexes_against_searchsorted():
    gc = cai.GaussianConditional(None)
    gc.update_scale_table(cai.get_scale_table())
    s = torch.exp(torch.empty(5000).uniform_(math.log(0.01), math.log(400)))
    s[:64] = gc.scale_table  # exact table hits: (scales <= s) is inclusive
    idx = gc.build_indexes(s.view(1, 1, 50, 100)).flatten()
    want = torch.searchsorted(gc.scale_table[:-1].contiguous(), s.clamp(min=0.11), right=False)
    assert torch.equal(idx.long(), want)
    assert idx.min().item() == 0 and idx.max().item() == 63


def test_entropy_bottleneck_pmf_sums_to_one():
    eb = cai.EntropyBottleneck(4).eval()
    with torch.no_grad():
        for f in range(4):
            getattr(eb, f"_factor{f}").normal_(0, 0.3)
        ks = torch.arange(-4000, 4001, dtype=torch.float32).view(1, 1, -1, 1).repeat(1, 4, 1, 1)
        _, lik = eb(ks)
    tot = lik.sum(dim=2).flatten()
    assert torch.allclose(tot, torch.ones_like(tot), atol=2e-3), tot
    assert {k for k in eb.state_dict()} >= {"_matrix0", "_bias4", "_factor3", "quantiles", "target", "_offset",
                                            "_quantized_cdf", "_cdf_length", "likelihood_lower_bound.bound"}


def test_hyperprior_forward_contract():
    m = cai.MeanScaleHyperprior(8, 12).eval()
    with torch.no_grad():
        out = m(torch.rand(1, 3, 64, 64))
    assert out["x_hat"].shape == (1, 3, 64, 64)
    assert out["likelihoods"]["y"].shape == (1, 12, 4, 4) and out["likelihoods"]["z"].shape == (1, 8, 1, 1)
