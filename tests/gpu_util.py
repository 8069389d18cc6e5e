import torch

BF16, F32 = torch.bfloat16, torch.float32


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def max_err(a, b) -> float:
    return float((a.float() - b.float()).abs().max())


def randn(*shape, scale=1.0, seed=0, dtype=BF16, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(device)
