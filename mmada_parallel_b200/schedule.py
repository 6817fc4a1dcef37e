"""Host-side integer/schedule logic of the generation loops (cheap, runs once per call or once per step on the CPU).

Mirrors, name for name, the helpers of the reference:
  cosine_schedule            A/generators/parallel_generator.py:73-75, M/models/sampling.py:39-40
  get_num_transfer_tokens    A/generators/parallel_generator.py:78-99 (floor form)
  get_num_transfer_tokens_m  M/models/modeling_mmada.py:63-81 (base + remainder form)
"""
from __future__ import annotations

import math
from typing import Callable, List

import torch


def cosine_schedule(t: torch.Tensor) -> torch.Tensor:
    return torch.cos(t * math.pi / 2)


def get_num_transfer_tokens(total_masks: int, text_steps: int) -> List[int]:
    """Per-step un-mask counts for one batch row (A)."""
    out, remaining = [], total_masks
    for step in range(text_steps):
        ratio = (step + 1) / text_steps
        target_remaining = int(total_masks * (1 - ratio))
        tokens_to_unmask = max(0, remaining - target_remaining)
        out.append(tokens_to_unmask)
        remaining -= tokens_to_unmask
    return out


def get_num_transfer_tokens_m(mask_num: int, steps: int) -> List[int]:
    base, remainder = mask_num // steps, mask_num % steps
    return [base + (1 if i < remainder else 0) for i in range(steps)]


def image_generation_step_indices(text_steps: int, timesteps: int) -> List[int]:
    """torch.linspace(text_steps // 4, text_steps - 1, timesteps).round().int()  (parallel_generator.py:157-159)."""
    return torch.linspace(text_steps // 4, text_steps - 1, timesteps).round().int().tolist()


def scheduled_mask_len(num_vq_tokens: int, step: int, text_steps: int, noise_schedule: Callable = cosine_schedule) -> int:
    """floor(num_vq_tokens * noise_schedule(ratio)) evaluated on a 0-d fp32 CPU tensor exactly like the reference
    (parallel_generator.py:318-322); at ratio == 1 the fp32 cosine is slightly negative and this returns -1."""
    ratio = 1.0 * (step + 1) / text_steps
    mask_ratio = noise_schedule(torch.tensor(ratio))
    return int((num_vq_tokens * mask_ratio).floor().item())


def stepwise_image_step_indices(text_steps: int) -> List[int]:
    """Image steps of the preview loop (A/app.py:162-164): 30 % of the steps, spread over the whole schedule."""
    return torch.linspace(0, text_steps - 1, int(text_steps * 0.3)).round().int().tolist()
