#!/usr/bin/env python
"""Which arithmetic does ATen CUDA use for tensor - python_scalar and tensor / python_scalar?"""
import numpy as np
import torch

torch.manual_seed(0)
x = torch.rand(1, 1, 192, 256, device="cuda")
for m, s in ((0.485, 0.229), (0.456, 0.224), (0.406, 0.225)):
    d = x - m
    print(f"m={m}: sub == x - float32(m) tensor: {(d == x - torch.tensor(np.float32(m), device='cuda')).float().mean().item():.4f}"
          f" | == (x.double() - m).float(): {(d == (x.double() - m).float()).float().mean().item():.4f}")
    q = d / s
    inv_f = np.float32(1.0) / np.float32(s)
    inv_d = np.float32(1.0 / s)
    print(f"  s={s}: inv_f={inv_f!r} inv_d={inv_d!r}")
    for name, alt in (("d * (1f/float(s))", d * torch.tensor(inv_f, device="cuda")),
                      ("d * float(1.0/s)", d * torch.tensor(inv_d, device="cuda")),
                      ("d / float(s) tensor", d / torch.tensor(np.float32(s), device="cuda")),
                      ("(d.double()/s).float()", (d.double() / s).float()),
                      ("(d.double()*(1/s)).float()", (d.double() * (1.0 / s)).float())):
        print(f"    {name:28s}: {(q == alt).float().mean().item():.4f}")
