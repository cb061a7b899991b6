"""Partitioning of tiles / sliding windows across the GPUs of one box.

Inference shards independent units (tiles, images, window rows of one scene) with NO data-path collective;
only the training-step config exchanges data (gradient all-reduce, emrt_b200/train.py)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n units for `rank` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_images(n_images: int, rank: int, world: int) -> List[int]:
    """Whole images round-robin to ranks, so all windows of an image are stitched locally (cfg 3)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_images, world))


def shard_scene_rows(row_origins: Sequence[int], crop_h: int, rank: int, world: int):
    """One big scene (cfg 5): contiguous bands of window ROWS per rank.  Returns (owned_rows, canvas_y0, canvas_y1,
    halo_rows): `owned_rows` are the window-row indices this rank stitches into label rows [canvas_y0, canvas_y1);
    `halo_rows` are neighbours' window rows that also cover those label rows and are RECOMPUTED locally
    (zero communication) so that every label row sums exactly the windows the reference would."""
    n = len(row_origins)
    b, e = shard_range(n, rank, world)
    if b == e:
        return [], 0, 0, []
    H = max(row_origins) + crop_h
    # label rows are split at the first row owned by each rank's first window row
    y0 = 0 if rank == 0 else row_origins[b]
    nb, ne = shard_range(n, rank + 1, world) if rank + 1 < world else (n, n)
    y1 = H if (rank + 1 >= world or nb == ne) else row_origins[nb]
    halo = [r for r in range(n) if not (b <= r < e) and row_origins[r] < y1 and row_origins[r] + crop_h > y0]
    return list(range(b, e)), y0, y1, halo
