"""Value/reward categorical transform helpers (muax/utils.py:65-102) on torch tensors, for host-side users
(tracers, learners).  The search kernels have their own bit-exact device versions (include/mz_math.h)."""
import torch


def _scaling(x, eps: float = 1e-3):  # muax/utils.py:65-67
    return torch.sign(x) * (torch.sqrt(torch.abs(x) + 1) - 1) + eps * x


def _inv_scaling(x, eps: float = 1e-3):  # muax/utils.py:70-76
    return torch.sign(x) * (((torch.sqrt(1 + 4 * eps * (torch.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1)


def scalar_to_support(x, support_size):  # muax/utils.py:79-91
    x = torch.clamp(_scaling(x), -support_size, support_size)
    low = torch.floor(x).to(torch.int64)
    high = torch.ceil(x).to(torch.int64)
    prob_high = x - low
    prob_low = 1.0 - prob_high
    n = 2 * support_size + 1
    lo = torch.nn.functional.one_hot(low + support_size, n) * prob_low[..., None]
    hi = torch.nn.functional.one_hot(high + support_size, n) * prob_high[..., None]
    return lo + hi


def support_to_scalar(probs, support_size):  # muax/utils.py:94-102
    rng = torch.arange(2 * support_size + 1, device=probs.device, dtype=probs.dtype) - support_size
    return _inv_scaling(torch.sum(rng * probs, dim=-1))
