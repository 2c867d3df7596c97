"""Window sharding for N > 1: windows are independent units, so rank r takes a contiguous range and one all-gather
collects the per-window results (binarised segmentations, embeddings) - SURVEY.md section 8(e)."""
from __future__ import annotations

from typing import Tuple

import torch


def window_range(num_windows: int, rank: int, world: int) -> Tuple[int, int, int]:
    """-> (first, one-past-last, windows-per-rank incl. padding) for `rank`."""
    per = (num_windows + world - 1) // world
    return min(rank * per, num_windows), min((rank + 1) * per, num_windows), per


def gather_windows(local: torch.Tensor, num_windows: int, world: int) -> torch.Tensor:
    """local: (per, ...) this rank's (zero padded) slice -> (num_windows, ...) on every rank.  One collective."""
    if world == 1:
        return local[:num_windows]
    import torch.distributed as dist
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out[:num_windows]
