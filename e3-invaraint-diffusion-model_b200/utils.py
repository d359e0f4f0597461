"""Host-side mirror of the types `sequence_model/utils.py` hands to the hot path: the discrete noise
schedule and the transition matrices.  They are tiny per-step table
builders that run on the host exactly as in the reference (quirk Q8: alphas_bar lives on the CPU); the
resulting [T,3,20,20] tables are uploaded once and consumed by the CUDA reverse-step kernel.
Same class names, constructor arguments, methods and error behaviour as the reference."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "blosum_substitute.npz")


def cosine_beta_schedule_discrete(timesteps, s=0.008):
    """reference utils.py:99-108."""
    steps = timesteps + 2
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(0.5 * np.pi * ((x / steps) + s) / (1 + s)) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    alphas = alphas_cumprod[1:] / alphas_cumprod[:-1]
    betas = 1 - alphas
    return betas.squeeze()


class PredefinedNoiseScheduleDiscrete(torch.nn.Module):
    """reference utils.py:206-233."""

    def __init__(self, noise_schedule, timesteps):
        super().__init__()
        self.timesteps = timesteps
        betas = cosine_beta_schedule_discrete(timesteps)
        self.register_buffer("betas", torch.from_numpy(betas).float())
        self.alphas = 1 - torch.clamp(self.betas, min=0, max=0.9999)
        log_alpha_bar = torch.cumsum(torch.log(self.alphas), dim=0)
        self.alphas_bar = torch.exp(log_alpha_bar)

    def forward(self, t_normalized=None, t_int=None):
        assert int(t_normalized is None) + int(t_int is None) == 1
        if t_int is None:
            t_int = torch.round(t_normalized * self.timesteps)
        return self.betas[t_int.long()]

    def get_alpha_bar(self, t_normalized=None, t_int=None):
        assert int(t_normalized is None) + int(t_int is None) == 1
        if t_int is None:
            t_int = torch.round(t_normalized * self.timesteps)
        return self.alphas_bar.to(t_int.device)[t_int.long()]


class DiscreteUniformTransition:
    """reference utils.py:235-271."""

    def __init__(self, x_classes: int):
        self.X_classes = x_classes
        self.u_x = torch.ones(1, self.X_classes, self.X_classes)
        if self.X_classes > 0:
            self.u_x = self.u_x / self.X_classes

    def get_Qt(self, beta_t, device):
        beta_t = beta_t.unsqueeze(1).to(device)
        self.u_x = self.u_x.to(device)
        return beta_t * self.u_x + (1 - beta_t) * torch.eye(self.X_classes, device=device).unsqueeze(0)

    def get_Qt_bar(self, alpha_bar_t, device):
        alpha_bar_t = alpha_bar_t.unsqueeze(1).to(device)
        self.u_x = self.u_x.to(device)
        return alpha_bar_t * torch.eye(self.X_classes, device=device).unsqueeze(0) + (1 - alpha_bar_t) * self.u_x


def _load_blosum(blosum_path):
    """Accepts the reference's own `blosum_substitute.pt` (utils.py:276) and falls back, like the
    reference's '../' retry (utils.py:277-279), to the copy of the three arrays shipped with this package."""
    for p in (blosum_path, "../" + blosum_path):
        if os.path.exists(p):
            d = torch.load(p)
            return d["original_score"], d["Qtb_temperature"], d["Qt_temperature"]
    if os.path.exists(_DATA):
        z = np.load(_DATA)
        return (torch.from_numpy(z["original_score"]), torch.from_numpy(z["Qtb_temperature"]),
                torch.from_numpy(z["Qt_temperature"]))
    raise FileNotFoundError(blosum_path)


class BlosumTransition:
    """reference utils.py:273-314 (temperature tables re-interpolated to timestep+1 entries; the
    shape test at :286 is always true)."""

    def __init__(self, blosum_path="./blosum_substitute.pt", x_classes=20, timestep=500):
        self.original_score, self.temperature_list, self.Qt_temperature = _load_blosum(blosum_path)
        self.X_classes = x_classes
        self.timestep = timestep
        t = self.temperature_list.float()[None, None]
        q = self.Qt_temperature.float()[None, None]
        self.temperature_list = F.interpolate(t, size=timestep + 1, mode="linear", align_corners=True).squeeze()
        self.Qt_temperature = F.interpolate(q, size=timestep + 1, mode="linear", align_corners=True).squeeze()

    def get_Qt_bar(self, t_normal, device):
        self.original_score = self.original_score.to(device)
        self.temperature_list = self.temperature_list.to(device)
        t_int = torch.round(t_normal * self.timestep).to(device)
        temperatue = self.temperature_list[t_int.long()]
        q_x = self.original_score.unsqueeze(0) / temperatue.unsqueeze(2)
        q_x = torch.softmax(q_x, dim=2)
        q_x[q_x < 1e-6] = 1e-6
        return q_x

    def get_Qt(self, t_normal, device):
        self.original_score = self.original_score.to(device)
        self.Qt_temperature = self.Qt_temperature.to(device)
        t_int = torch.round(t_normal * self.timestep).to(device)
        temperatue = self.Qt_temperature[t_int.long()]
        q_x = self.original_score.unsqueeze(0) / temperatue.unsqueeze(2)
        return torch.softmax(q_x, dim=2)


def step_tables(t, s, noise_schedule, transition):
    """(Qt, Qsb, Qtb) of reference sample.py:156-160, computed on the HOST with the caller's own
    schedule/transition objects (duck-typed) -> float32 [n, 3, 20, 20]."""
    cpu = torch.device("cpu")
    alpha_t_bar = noise_schedule.get_alpha_bar(t_normalized=t.cpu())
    alpha_s_bar = noise_schedule.get_alpha_bar(t_normalized=s.cpu())
    Qtb = transition.get_Qt_bar(alpha_t_bar, cpu)
    Qsb = transition.get_Qt_bar(alpha_s_bar, cpu)
    Qt = (Qsb / Qtb) / (Qsb / Qtb).sum(dim=-1).unsqueeze(dim=2)
    return torch.stack([Qt, Qsb, Qtb], dim=1).float().contiguous()


def loop_tables(timesteps, noise_schedule, transition):
    """Tables for every step of denoise() (reference sample.py:192-197): entry s_int holds the triple
    for s = s_int/T, t = (s_int+1)/T, built with the reference's own float32 arithmetic."""
    s_array = torch.arange(timesteps, dtype=torch.float32).unsqueeze(1) * torch.ones((1, 1))
    s_norm = s_array / timesteps
    t_norm = (s_array + 1) / timesteps
    return step_tables(t_norm, s_norm, noise_schedule, transition)


def elbo_loss(logits1, logits2, eps=1e-6):
    """reference utils.py:132-161 (training-side; plain torch ops on the logits the CUDA forward produced)."""
    probs1 = F.softmax(logits1, dim=-1)
    probs2 = F.softmax(logits2, dim=-1)
    log_probs1 = F.log_softmax(logits1 + eps, dim=-1)
    kl_div = F.kl_div(log_probs1, probs2, reduction="batchmean")
    nll = -torch.mean(torch.sum(probs1 * log_probs1, dim=-1))
    return nll + kl_div
