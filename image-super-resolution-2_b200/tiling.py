"""Overlapped-tile scheduler: the index math of the reference's `_tiled_forward`
(models/team29_FreqFusion/io.py:82-121) plus batched execution and the stitch kernel.

Tile origins, blend ramps and the accumulation order are identical to the reference; the only change is
that tiles are run through the model in batches instead of one at a time.
"""
import ctypes as C_

import torch

from . import lib as L


def tile_positions(size, tile, overlap):
    """io.py:88-95 -- origins along one axis: range(0, max(size-tile+1, 1), tile-overlap) plus a flush-right tail."""
    step = tile - overlap
    pos = list(range(0, max(size - tile + 1, 1), step))
    if pos[-1] + tile < size:
        pos.append(size - tile)
    return pos


def axis_weights(positions, size, tile, overlap, scale=4):
    """io.py:104-116 -- 1-D blend weights per tile origin: linspace ramps of length min(overlap*scale, tile*scale//4)
    on the sides that have a neighbour, 1 elsewhere.  Returns fp32 [len(positions), tile*scale]."""
    st = tile * scale
    blend = min(overlap * scale, st // 4)
    out = torch.ones(len(positions), st, dtype=torch.float32)
    if blend > 0:
        ramp = torch.linspace(0, 1, blend)
        for i, p in enumerate(positions):
            if p > 0:
                out[i, :blend] = ramp
            if p + tile < size:
                out[i, -blend:] = 1 - ramp
    return out


def plan(h, w, tile, overlap, scale=4):
    if h < tile or w < tile:
        raise ValueError(f"image {h}x{w} is smaller than the tile {tile} (the reference's tile path fails here too)")
    ys, xs = tile_positions(h, tile, overlap), tile_positions(w, tile, overlap)
    return dict(ys=ys, xs=xs, wy=axis_weights(ys, h, tile, overlap, scale), wx=axis_weights(xs, w, tile, overlap, scale),
                tile=tile, scale=scale, h=h, w=w)


def choose_tile(h, w):
    """io.main uses tile 128 / overlap 32 (io.py:226); images with a side below 128 use the function defaults 64 / 8."""
    if min(h, w) >= 128:
        return 128, 32
    if min(h, w) >= 64:
        return 64, 8
    raise ValueError(f"image {h}x{w}: sides below 64 px are not supported by the tile scheduler")


def extract_tiles(lr, pl):
    """lr: [1,3,h,w] -> [T,3,tile,tile] in (y-major, x-minor) order."""
    t = pl["tile"]
    return torch.stack([lr[0, :, y:y + t, x:x + t] for y in pl["ys"] for x in pl["xs"]]).contiguous()


class Stitcher:
    """Device-side state of one plan (origins, weights) + the ff_stitch call."""

    def __init__(self, pl, device):
        self.pl = pl
        s = pl["scale"]
        self.ty = torch.tensor([y * s for y in pl["ys"]], dtype=torch.int32, device=device)
        self.tx = torch.tensor([x * s for x in pl["xs"]], dtype=torch.int32, device=device)
        self.wy = pl["wy"].to(device).contiguous()
        self.wx = pl["wx"].to(device).contiguous()
        self.H, self.W = pl["h"] * s, pl["w"] * s

    def __call__(self, tiles, out=None, out_u8=None):
        """tiles: fp32 [T,3,ts,ts] on the device -> out fp32 [3,H,W] and/or out_u8 uint8 [H,W,3]."""
        pl = self.pl
        ts = pl["tile"] * pl["scale"]
        assert tiles.is_cuda and tiles.dtype == torch.float32 and tiles.is_contiguous()
        assert tiles.shape == (len(pl["ys"]) * len(pl["xs"]), 3, ts, ts)
        L.check(L.load().ff_stitch(C_.c_void_p(tiles.data_ptr()), C_.c_void_p(self.ty.data_ptr()), C_.c_void_p(self.tx.data_ptr()),
                                   C_.c_void_p(self.wy.data_ptr()), C_.c_void_p(self.wx.data_ptr()), len(pl["ys"]), len(pl["xs"]), ts,
                                   self.H, self.W, C_.c_void_p(out.data_ptr()) if out is not None else None,
                                   C_.c_void_p(out_u8.data_ptr()) if out_u8 is not None else None,
                                   C_.c_void_p(torch.cuda.current_stream().cuda_stream)), "ff_stitch")
        return out if out is not None else out_u8
