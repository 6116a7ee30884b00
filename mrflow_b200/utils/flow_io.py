"""``utils/flow_io.py`` of the reference (the .flo half) on the B200.

``flow_read(filename)`` keeps the reference's name, argument and ``(u, v)`` return (:33-48) and its two assertions
(tag, size); the file goes to HBM as it lies on disk and is split into the two planes there
(``mrf_flo_decode``).  ``device=True`` returns the fp32 CUDA tensors instead of numpy arrays, which is what the
path's device-resident entry points (``planeparallax.Band``) take.  ``flow_write`` writes a file the reference's
``flow_read`` accepts (the reference's own writer fails on Python 3: it writes a ``str`` tag to a binary file, :78).
"""
import numpy as np
import torch

from .. import device
from .._host import require_cuda

TAG_FLOAT = 202021.25          # :28
TAG_CHAR = b"PIEH"             # :29


def flow_read(filename, device_out=False):
    """utils/flow_io.py:33-48.  Returns ``(u, v)``: fp32 [h,w] numpy arrays, or CUDA tensors with ``device_out``."""
    require_cuda()
    raw = np.fromfile(filename, dtype=np.uint8)
    assert raw.size >= 12, " flow_read:: file too short"
    check = raw[:4].view(np.float32)[0]
    assert check == TAG_FLOAT, " flow_read:: Wrong tag in flow file (should be: {0}, is: {1}). Big-endian machine? ".format(
        TAG_FLOAT, check)
    width, height = int(raw[4:8].view(np.int32)[0]), int(raw[8:12].view(np.int32)[0])
    size = width * height
    assert width > 0 and height > 0 and size > 1 and size < 100000000, \
        " flow_read:: Wrong input size (width = {0}, height = {1}).".format(width, height)
    assert raw.size >= 12 + 8 * size, " flow_read:: file shorter than its header says"      # numpy's reshape would raise
    staged = torch.from_numpy(raw).pin_memory()
    u, v = device.flo_decode(staged.cuda(non_blocking=True), height, width)
    if device_out:
        return u, v
    return u.cpu().numpy(), v.cpu().numpy()


def flow_write(filename, uv, v=None):
    """utils/flow_io.py:50-100 (same layout; the tag is written as bytes)."""
    if v is None:
        uv_ = np.asarray(uv)
        assert uv_.ndim == 3
        if uv_.shape[0] == 2:
            u, v = uv_[0], uv_[1]
        elif uv_.shape[2] == 2:
            u, v = uv_[:, :, 0], uv_[:, :, 1]
        else:
            raise ValueError("Wrong format for flow input")
    else:
        u = np.asarray(uv)
        v = np.asarray(v)
    assert u.shape == v.shape
    height, width = u.shape
    with open(filename, "wb") as f:
        f.write(TAG_CHAR)
        np.array(width).astype(np.int32).tofile(f)
        np.array(height).astype(np.int32).tofile(f)
        tmp = np.zeros((height, width * 2), dtype=np.float32)
        tmp[:, 0::2] = u
        tmp[:, 1::2] = v
        tmp.tofile(f)
