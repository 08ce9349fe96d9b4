"""ORACLE (test infrastructure only).  CPU restatement of `_tiled_forward` and the image quantisation of
/root/reference/models/team29_FreqFusion/io.py:71-76, 82-121 for any `model` callable."""
import torch


def tiled_forward(model, lr_img, tile_size=64, overlap=8, scale=4):
    _, _, h, w = lr_img.shape
    sr = torch.zeros(1, 3, h * scale, w * scale)
    wm = torch.zeros(1, 1, h * scale, w * scale)
    step = tile_size - overlap
    ys = list(range(0, max(h - tile_size + 1, 1), step))
    if ys[-1] + tile_size < h:
        ys.append(h - tile_size)
    xs = list(range(0, max(w - tile_size + 1, 1), step))
    if xs[-1] + tile_size < w:
        xs.append(w - tile_size)
    st = tile_size * scale
    blend = min(overlap * scale, st // 4)
    for y in ys:
        for x in xs:
            tile = model(lr_img[:, :, y:y + tile_size, x:x + tile_size])
            wy, wx = torch.ones(st), torch.ones(st)
            if blend > 0:
                ramp = torch.linspace(0, 1, blend)
                if y > 0:
                    wy[:blend] = ramp
                if y + tile_size < h:
                    wy[-blend:] = 1 - ramp
                if x > 0:
                    wx[:blend] = ramp
                if x + tile_size < w:
                    wx[-blend:] = 1 - ramp
            wgt = (wy.unsqueeze(1) * wx.unsqueeze(0)).unsqueeze(0).unsqueeze(0)
            sr[:, :, y * scale:y * scale + st, x * scale:x * scale + st] += tile * wgt
            wm[:, :, y * scale:y * scale + st, x * scale:x * scale + st] += wgt
    return sr / wm.clamp(min=1e-8), ys, xs


def to_uint8(t):
    """io._save_image: clamp(0,1)*255 -> np.round (half to even) -> uint8, HWC."""
    if t.dim() == 4:
        t = t.squeeze(0)
    return (t.clamp(0, 1).permute(1, 2, 0).numpy() * 255.0).round().astype("uint8")
