"""Multi-GPU work sharding for the tile scheduler: the path is embarrassingly parallel, so ranks
get disjoint sets of images (or, for a single large image, disjoint tile ranges) and no data-path collective exists.
The only exchange is a final gather of small per-image records (names / checksums / timings) to rank 0.
"""
import torch.distributed as dist

from . import tiling


def tile_count(h, w):
    t, ov = tiling.choose_tile(h, w)
    return len(tiling.tile_positions(h, t, ov)) * len(tiling.tile_positions(w, t, ov))


def assign_images(costs, world):
    """Longest-processing-time-first assignment of images to ranks.  costs: per-image tile counts.
    Returns a list (len world) of sorted image-index lists; deterministic (ties broken by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


def assign_tiles(num_tiles, world):
    """Contiguous split of one image's tile list (y-major order) over ranks: rank r gets [lo, hi)."""
    base, rem = divmod(num_tiles, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def gather_records(local_records):
    """Final gather (the only collective): every rank contributes {image_index: record}; rank 0 gets the merged dict."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local_records)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, dict(local_records))
    merged = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError(f"image {k} was processed by two ranks")
            merged[k] = v
    return merged
