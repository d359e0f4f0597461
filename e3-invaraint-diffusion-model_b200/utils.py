"""Host-side table builders behind the `sequence_model/utils.py` types of the reference.

What the hot path needs from the reference's schedule / transition objects is a handful of tiny tables: alpha-bar per step,
BLOSUM temperatures per step, and from them the (Qt, Qsb, Qtb) triple of every reverse step.  They are built ONCE on the host
in fp32 with the arithmetic the reference uses on the CPU (quirk Q8: alphas_bar lives on the CPU there too), uploaded as a
[T,3,20,20] array and consumed by the CUDA reverse-step kernel.  The classes below keep the reference's names, constructor
arguments, method signatures and error behaviour (they are the drop-in boundary), but they are thin fronts of three
vectorised builders -- `qbar_blosum`, `qbar_uniform`, `posterior_tables` -- and never move state between devices.

The order of the fp32 / fp64 operations is dictated by the parity requirement (tables must be bit-identical to the
reference's: tests/test_oracle_pin.py::test_product_host_tables_match_oracle); everything else is this package's own.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "blosum_substitute.npz")
_CPU = torch.device("cpu")


# ---------------------------------------------------------------------------------------------------------------------
# cosine schedule (reference utils.py:99-108, 206-233)
# ---------------------------------------------------------------------------------------------------------------------
def cosine_beta_schedule_discrete(timesteps: int, s: float = 0.008) -> np.ndarray:
    """betas[0..T] (float64) of the cosine schedule on the T+2 grid points the reference samples."""
    n = timesteps + 2
    grid = np.linspace(0, n, n)
    f = np.cos(0.5 * np.pi * ((grid / n) + s) / (1 + s)) ** 2
    f = f / f[0]
    return (1 - f[1:] / f[:-1]).squeeze()


def _alpha_bar_table(timesteps: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(betas, alphas, alphas_bar), fp32, T+1 entries: alphas_bar = exp(cumsum(log(1 - clamp(beta, 0, 0.9999))))."""
    betas = torch.from_numpy(cosine_beta_schedule_discrete(timesteps)).float()
    alphas = 1 - betas.clamp(min=0, max=0.9999)
    return betas, alphas, alphas.log().cumsum(0).exp()


def _step_index(t_normalized, t_int, timesteps: int) -> torch.Tensor:
    # the reference's guard (utils.py:224,230): exactly one of the two must be given
    assert int(t_normalized is None) + int(t_int is None) == 1
    if t_int is None:
        t_int = torch.round(t_normalized * timesteps)
    return t_int.long()


class PredefinedNoiseScheduleDiscrete(torch.nn.Module):
    """Lookup tables of the discrete noise schedule.  `betas` is a registered buffer (it is part of the reference
    checkpoint: key `discrete_noise_schedule.betas`); `alphas` / `alphas_bar` are plain CPU attributes."""

    def __init__(self, noise_schedule, timesteps):
        super().__init__()
        self.timesteps = timesteps
        betas, self.alphas, self.alphas_bar = _alpha_bar_table(timesteps)
        self.register_buffer("betas", betas)

    def forward(self, t_normalized=None, t_int=None):
        return self.betas[_step_index(t_normalized, t_int, self.timesteps)]

    def get_alpha_bar(self, t_normalized=None, t_int=None):
        idx = _step_index(t_normalized, t_int, self.timesteps)
        return self.alphas_bar.to(idx.device)[idx]


# ---------------------------------------------------------------------------------------------------------------------
# transition matrices (reference utils.py:235-314)
# ---------------------------------------------------------------------------------------------------------------------
def qbar_uniform(stay: torch.Tensor, spread: torch.Tensor, classes: int) -> torch.Tensor:
    """[n,C,C] = stay_n * I + spread_n / C for n scalar pairs (any shape with n elements)."""
    eye = torch.eye(classes).unsqueeze(0)
    flat = torch.ones(1, classes, classes) / max(classes, 1)
    return stay.reshape(-1, 1, 1) * eye + spread.reshape(-1, 1, 1) * flat


def qbar_blosum(score: torch.Tensor, temperatures: torch.Tensor, clamp: Optional[float]) -> torch.Tensor:
    """[n,C,C] row-softmax of score / temperature_n; entries below `clamp` are raised to it (quirk Q7: rows then no
    longer sum to one -- the reference does the same)."""
    q = torch.softmax(score.unsqueeze(0) / temperatures.reshape(-1, 1, 1), dim=2)
    return q if clamp is None else torch.where(q < clamp, torch.full_like(q, clamp), q)


class DiscreteUniformTransition:
    def __init__(self, x_classes: int):
        self.X_classes = x_classes
        self.u_x = torch.ones(1, x_classes, x_classes) / max(x_classes, 1)

    def get_Qt(self, beta_t, device):
        b = beta_t.detach().to(_CPU)
        return qbar_uniform(1 - b, b, self.X_classes).to(device)

    def get_Qt_bar(self, alpha_bar_t, device):
        a = alpha_bar_t.detach().to(_CPU)
        return qbar_uniform(a, 1 - a, self.X_classes).to(device)


def _blosum_arrays(blosum_path):
    """(score [20,20], Qtb temperatures [500], Qt temperatures [500]).  The reference's own `blosum_substitute.pt` is used when
    it is found at `blosum_path` or one directory up (the reference's retry, utils.py:277-279); otherwise the copy of the three
    arrays that ships with this package."""
    for cand in (blosum_path, os.path.join("..", blosum_path)):
        if cand and os.path.exists(cand):
            blob = torch.load(cand)
            return blob["original_score"], blob["Qtb_temperature"], blob["Qt_temperature"]
    if not os.path.exists(_DATA):
        raise FileNotFoundError(blosum_path)
    z = np.load(_DATA)
    return tuple(torch.from_numpy(z[k]) for k in ("original_score", "Qtb_temperature", "Qt_temperature"))


_load_blosum = _blosum_arrays  # (name kept for callers of the previous round)


def _resample(table: torch.Tensor, n: int) -> torch.Tensor:
    """500-entry temperature table -> n entries, linear with aligned end points."""
    return F.interpolate(table.float()[None, None], size=n, mode="linear", align_corners=True).squeeze()


class BlosumTransition:
    """BLOSUM-substitution transition: Q(x) = softmax(score / temperature[round(x * timestep)]).  The callers hand it
    alpha-bar where the table was built for normalised time (quirk Q2) -- reproduced as is."""

    def __init__(self, blosum_path="./blosum_substitute.pt", x_classes=20, timestep=500):
        score, qtb_temp, qt_temp = _blosum_arrays(blosum_path)
        self.original_score = score
        self.X_classes = x_classes
        self.timestep = timestep
        self.temperature_list = _resample(qtb_temp, timestep + 1)
        self.Qt_temperature = _resample(qt_temp, timestep + 1)

    def _rows(self, table, t_normal):
        return table[torch.round(t_normal.detach().to(_CPU) * self.timestep).long()]

    def get_Qt_bar(self, t_normal, device):
        return qbar_blosum(self.original_score, self._rows(self.temperature_list, t_normal), 1e-6).to(device)

    def get_Qt(self, t_normal, device):
        return qbar_blosum(self.original_score, self._rows(self.Qt_temperature, t_normal), None).to(device)


# ---------------------------------------------------------------------------------------------------------------------
# per-step posterior tables for the CUDA reverse step
# ---------------------------------------------------------------------------------------------------------------------
def posterior_tables(alpha_bar_t: torch.Tensor, alpha_bar_s: torch.Tensor, transition) -> torch.Tensor:
    """fp32 [n,3,C,C] = (Qt, Qsb, Qtb) for n (t, s) pairs: Qtb / Qsb from the caller's (duck-typed) transition object,
    Qt = row-normalised Qsb / Qtb (the reference's stand-in for the one-step matrix, sample.py:160, quirk Q5)."""
    Qtb = transition.get_Qt_bar(alpha_bar_t, _CPU)
    Qsb = transition.get_Qt_bar(alpha_bar_s, _CPU)
    ratio = Qsb / Qtb
    Qt = ratio / ratio.sum(dim=-1).unsqueeze(dim=2)
    return torch.stack([Qt, Qsb, Qtb], dim=1).float().contiguous()


def step_tables(t, s, noise_schedule, transition):
    """Tables of one reverse step for per-graph normalised times t, s [B,1] (reference sample.py:156-160)."""
    return posterior_tables(noise_schedule.get_alpha_bar(t_normalized=t.cpu()), noise_schedule.get_alpha_bar(t_normalized=s.cpu()), transition)


def loop_tables(timesteps, noise_schedule, transition):
    """Tables of every step of a T-step sampling: entry k serves the step s_int = k, i.e. s = k/T, t = (k+1)/T, formed in
    fp32 exactly like the loop variables of reference sample.py:192-197."""
    k = torch.arange(timesteps, dtype=torch.float32).unsqueeze(1) * torch.ones((1, 1))
    return step_tables((k + 1) / timesteps, k / timesteps, noise_schedule, transition)
