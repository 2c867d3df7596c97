"""Window sharding for N > 1: windows are independent units, so rank r takes a contiguous range and one all-gather
collects the per-window results (binarised segmentations, embeddings) - SURVEY.md section 8(e)."""
from __future__ import annotations

from typing import Tuple

import torch


def window_range(num_windows: int, rank: int, world: int) -> Tuple[int, int, int]:
    """-> (first, one-past-last, windows-per-rank incl. padding) for `rank`."""
    per = (num_windows + world - 1) // world
    return min(rank * per, num_windows), min((rank + 1) * per, num_windows), per


def window_ranges(num_windows: int, world: int, root: int = 0, root_share: float = None):
    """-> ([(first, one-past-last)] per rank, rows per rank incl. padding).  `root_share` in (0, 1]: the rank that also
    clusters (`root`) takes that fraction of an even share and the others split the rest, so that - with recordings processed
    back to back - its networks + clustering take as long as the other ranks' networks alone."""
    if world == 1 or not root_share or root_share >= 1.0:
        per = (num_windows + world - 1) // world
        return [(min(r * per, num_windows), min((r + 1) * per, num_windows)) for r in range(world)], per
    n_root = max(0, min(num_windows, int(round(root_share * num_windows / world))))
    rest = num_windows - n_root
    others = world - 1
    base, extra = divmod(rest, others)
    out, pos, k = [], 0, 0
    for r in range(world):
        n = n_root if r == root else base + (1 if k < extra else 0)
        if r != root:
            k += 1
        out.append((pos, pos + n))
        pos += n
    return out, max(1, max(b - a for a, b in out))


def gather_windows(local: torch.Tensor, num_windows: int, world: int) -> torch.Tensor:
    """local: (per, ...) this rank's (zero padded) slice -> (num_windows, ...) on every rank.  One collective."""
    if world == 1:
        return local[:num_windows]
    import torch.distributed as dist
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out[:num_windows]


def pack_records(seg: torch.Tensor, stats: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """(per,T,S) uint8 + (per,S,2) int32 + (per,S,D) fp32 -> (per, record_bytes) uint8, 4-byte aligned fields."""
    per = seg.shape[0]
    nseg = seg[0].numel()
    o1 = (nseg + 3) // 4 * 4
    o2 = o1 + stats[0].numel() * 4
    rec = o2 + emb[0].numel() * 4
    buf = torch.zeros((per, rec), dtype=torch.uint8, device=seg.device)
    buf[:, :nseg] = seg.reshape(per, nseg)
    buf[:, o1:o2] = stats.contiguous().view(torch.uint8).reshape(per, -1)
    buf[:, o2:] = emb.contiguous().view(torch.uint8).reshape(per, -1)
    return buf


def unpack_records(buf: torch.Tensor, seg_shape, stats_shape, emb_shape):
    """inverse of pack_records for `buf` (n, record_bytes); shapes are per-window shapes."""
    n = buf.shape[0]
    nseg = 1
    for d in seg_shape:
        nseg *= d
    nst = 1
    for d in stats_shape:
        nst *= d
    o1 = (nseg + 3) // 4 * 4
    o2 = o1 + nst * 4
    seg = buf[:, :nseg].contiguous().reshape((n,) + tuple(seg_shape))
    stats = buf[:, o1:o2].contiguous().view(torch.int32).reshape((n,) + tuple(stats_shape))
    emb = buf[:, o2:].contiguous().view(torch.float32).reshape((n,) + tuple(emb_shape))
    return seg, stats, emb


def gather_records(seg: torch.Tensor, stats: torch.Tensor, emb: torch.Tensor, num_windows: int, world: int, ranges=None):
    """The single data-path collective of the window-sharded mode: every rank contributes its packed per-window records
    (binarised segmentations, frame counters, embeddings); -> the three tensors for all `num_windows` windows.
    `ranges`: per-rank (first, one-past-last) when the split is uneven (window_ranges with a root share); every rank still
    sends the same number of (padded) rows, and the valid prefix of each rank's block is kept."""
    packed = pack_records(seg, stats, emb)
    if ranges is None or world == 1:
        buf = gather_windows(packed, num_windows, world)
    else:
        import torch.distributed as dist
        per = packed.shape[0]
        full = torch.empty((world * per, packed.shape[1]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(full, packed.contiguous())
        buf = torch.cat([full[r * per:r * per + (b - a)] for r, (a, b) in enumerate(ranges)])
    return unpack_records(buf, seg.shape[1:], stats.shape[1:], emb.shape[1:])
