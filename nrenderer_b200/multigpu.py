"""Sample-slice partitioning of one frame across the GPUs of a box (partitioning A).

One process per GPU (torchrun).  The scene (< 1 MB) is replicated; rank g renders samples
[g*spp/G, (g+1)*spp/G) of EVERY pixel into a linear fp32 accumulation buffer (rgb = sums, a = sample
count); the partial frames are combined with ONE reduce(sum) to rank 0 — NCCL over NVLink on GPUs, gloo
in the CPU tests — and only then resolved (÷ count, sqrt gamma: the reference applies the gamma to the
per-pixel MEAN, AccPathTracer.cpp:32-34, so the reduce must happen in linear space).  Because the
counter-based RNG is keyed by the GLOBAL sample index, the union of the slices is the same set of paths
a single GPU renders; only the fp32 summation order differs.

Host plumbing only: the arithmetic is `render_slice` (the C ABI's nrcu_render_accumulate on a GPU) and
`resolve` (nrcu_resolve).  Both are passed in so that the partition/reduce logic can be exercised on CPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def sample_slice(spp: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [s0, s1) of the sample indices [0, spp) owned by `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return rank * spp // world, (rank + 1) * spp // world


def all_slices(spp: int, world: int):
    return [sample_slice(spp, r, world) for r in range(world)]


def render_frame(render_slice: Callable[[torch.Tensor, int, int], Optional[dict]],
                 resolve: Callable[[torch.Tensor, torch.Tensor], None],
                 accum: torch.Tensor, rgba: Optional[torch.Tensor], spp: int,
                 rank: int = 0, world: int = 1, group=None, root: int = 0, collective_events=None) -> Optional[dict]:
    """Render this rank's sample slice into `accum` (zeroed here), reduce to `root`, resolve there.

    render_slice(accum, s0, s1) adds the linear sums of samples [s0, s1) to accum[..., :3] and the
    sample count to accum[..., 3]; an empty slice (more ranks than samples) must add nothing.
    resolve(accum, rgba) writes sqrt(accum.rgb / accum.a), alpha 1.  Returns render_slice's stats.

    STREAM CONTRACT: `accum.zero_()` and the reduce are issued on torch's CURRENT stream, so render_slice and resolve
    must run on that stream too - on a GPU: `ctx.set_stream(torch.cuda.current_stream().cuda_stream)` before the
    first frame (bench.py does) - or synchronise themselves (`ctx.synchronize()`) before returning.  A context left
    on its own non-blocking stream would race with the zeroing and with the collective.

    collective_events: optional (start, end) torch.cuda.Event pair recorded around the exchange step (reduce + resolve),
    so that the caller can report the collective's own device time.
    """
    s0, s1 = sample_slice(spp, rank, world)
    accum.zero_()
    stats = render_slice(accum, s0, s1) if s1 > s0 else None
    if collective_events is not None:
        collective_events[0].record()
    if world > 1:
        dist.reduce(accum, dst=root, op=dist.ReduceOp.SUM, group=group)
    if rank == root and rgba is not None:
        resolve(accum, rgba)
    if collective_events is not None:
        collective_events[1].record()
    return stats
