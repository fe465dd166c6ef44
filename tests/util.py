"""Shared helpers for the parity tests."""
import torch


def rel_err(a, b):
    """max |a-b| / max |b| — the relative error the parity bar (BASELINE.json north_star: 1e-2 bf16) is stated in."""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def rel_l2(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def outlier_frac(a, b, tol=1e-2):
    """fraction of elements whose error exceeds tol * max|b| (ReLU-mask flips show up as a few full-magnitude
    outliers while the bulk agrees)"""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).abs() > tol * b.abs().max().clamp_min(1e-12)).float().mean().item()
