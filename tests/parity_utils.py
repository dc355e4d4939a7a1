import torch


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def to_dev(sd, device="cuda"):
    return {k: v.to(device) for k, v in sd.items()}


# north_star tolerance for the bf16 encoder / Q-Former outputs: max|a-b|/max|b| <= 2e-2 and cosine >= 0.999
TOL_REL = 2e-2
TOL_COS = 0.999


def assert_parity(got, want, tag="", rel=TOL_REL, cos=TOL_COS):
    r, c = relerr(got, want), cosine(got, want)
    assert r <= rel and c >= cos, f"{tag}: relerr {r:.3e} (tol {rel}), cosine {c:.6f} (tol {cos})"
    return r, c
