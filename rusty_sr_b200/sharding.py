"""Multi-GPU sharding plans for the forward path; one process per GPU.

* image batches: independent units, image i -> rank i mod world, no data-path collective;
* one large frame: horizontal bands with `HALO` rows of the 3-channel INPUT on each side
  ("overlap-recompute"): every rank runs the unmodified pipeline on its band + halo and keeps
  the centre rows.  HALO = 7 is the receptive field of sr_net (2+2+1+1+1 rows: conv0 5x5,
  conv1-3 5x5, then three 3x3 levels; network.rs:33,60-72).

The communication (scatter of input rows, gather of output rows) uses torch.distributed —
NCCL on GPUs, gloo in the CPU tests — and is the only place this package touches torch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

HALO = 7


@dataclass(frozen=True)
class Band:
    rank: int
    y0: int      # first output-owning LR row
    y1: int      # one past the last
    top: int     # halo rows available above (min(HALO, y0))
    bottom: int  # halo rows available below

    @property
    def rows_in(self) -> slice:
        return slice(self.y0 - self.top, self.y1 + self.bottom)


def band_plan(height: int, world: int, halo: int = HALO) -> List[Band]:
    """Splits `height` rows into `world` contiguous bands (sizes differ by at most one row).
    Ranks beyond the number of rows get empty bands."""
    bands = []
    for r in range(world):
        y0, y1 = r * height // world, (r + 1) * height // world
        bands.append(Band(r, y0, y1, min(halo, y0), min(halo, height - y1)))
    return bands


def batch_plan(n_images: int, world: int) -> List[List[int]]:
    """image i -> rank i mod world."""
    return [list(range(r, n_images, world)) for r in range(world)]


def upscale_frame_banded(forward_band: Callable[[np.ndarray, int, int], np.ndarray], frame: Optional[np.ndarray],
                         height: int, width: int, factor: int = 3, group=None, device=None) -> Optional[np.ndarray]:
    """Upscales one [H][W][3] f32 frame across all ranks of `group`.

    Rank 0 holds `frame`; every rank receives its band (+halo) of input rows, calls
    forward_band(rows, top_halo, bottom_halo) -> [(y1-y0)*f][W*f][3], and rank 0 gathers the
    output rows.  Returns the full output on rank 0, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    plan = band_plan(height, world)
    me = plan[rank]
    dev = device if device is not None else torch.device("cpu")
    # scatter input rows (each band is a different size: point-to-point from rank 0)
    if rank == 0:
        reqs = []
        for b in plan[1:]:
            if b.y1 > b.y0:
                t = torch.from_numpy(np.ascontiguousarray(frame[b.rows_in])).to(dev)
                reqs.append(dist.isend(t, dst=b.rank, group=group))
        rows = np.ascontiguousarray(frame[me.rows_in])
        for q in reqs:
            q.wait()
    else:
        n_in = me.y1 - me.y0 + me.top + me.bottom
        if me.y1 > me.y0:
            t = torch.empty((n_in, width, 3), dtype=torch.float32, device=dev)
            dist.recv(t, src=0, group=group)
            rows = t.cpu().numpy()
        else:
            rows = None
    out_band = forward_band(rows, me.top, me.bottom) if rows is not None and me.y1 > me.y0 else None
    # gather output rows on rank 0
    if rank == 0:
        out = np.empty((height * factor, width * factor, 3), np.float32)
        if out_band is not None:
            out[me.y0 * factor:me.y1 * factor] = out_band
        for b in plan[1:]:
            if b.y1 > b.y0:
                t = torch.empty(((b.y1 - b.y0) * factor, width * factor, 3), dtype=torch.float32, device=dev)
                dist.recv(t, src=b.rank, group=group)
                out[b.y0 * factor:b.y1 * factor] = t.cpu().numpy()
        return out
    if out_band is not None:
        dist.send(torch.from_numpy(np.ascontiguousarray(out_band)).to(dev), dst=0, group=group)
    return None
