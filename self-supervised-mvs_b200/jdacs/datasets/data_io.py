"""Mirror of `datasets/data_io.py` (jdacs/datasets/data_io.py:15-80; the same two functions live in jdacs-ms/dataset/data_io.py,
dataset/utils.py, test.py:120-155 and both fusion/depthfusion.py): the PFM reader / writer of the depth and confidence maps.

File format (kept byte for byte): header "Pf\\n" (grey) or "PF\\n" (colour), "<width> <height>\\n", "<scale>\\n" with a negative
scale for little-endian data ("%f" formatting), then the rows BOTTOM-UP as raw float32.  Host-side I/O like the reference's;
`save_pfm_flipped` takes a map whose rows are already bottom-up (ops.upsample_nearest(..., flip_rows=True) writes them so on the
GPU), which removes the host-side flip copy of a 1600x1200 map."""
from __future__ import annotations

import re
import sys

import numpy as np


def read_pfm(filename):
    """-> (data [H,W] or [H,W,3] float32 with the rows top-down, scale).  jdacs/datasets/data_io.py:15-50."""
    with open(filename, "rb") as f:
        magic = f.readline().decode("utf-8").rstrip()
        if magic == "PF":
            channels = 3
        elif magic == "Pf":
            channels = 1
        else:
            raise Exception("Not a PFM file.")
        dims = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not dims:
            raise Exception("Malformed PFM header.")
        width, height = int(dims.group(1)), int(dims.group(2))
        scale = float(f.readline().rstrip())
        order = "<" if scale < 0 else ">"
        body = np.fromfile(f, order + "f")
    shape = (height, width, 3) if channels == 3 else (height, width)
    return np.flipud(body.reshape(shape)), abs(scale)


def _header(image, scale):
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        magic = b"PF\n"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        magic = b"Pf\n"
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    little = image.dtype.byteorder == "<" or (image.dtype.byteorder == "=" and sys.byteorder == "little")
    return magic + ("%d %d\n" % (image.shape[1], image.shape[0])).encode("utf-8") + ("%f\n" % (-scale if little else scale)).encode("utf-8")


def save_pfm(filename, image, scale=1):
    """image [H,W] / [H,W,1] / [H,W,3] float32, rows top-down.  jdacs/datasets/data_io.py:53-80."""
    save_pfm_flipped(filename, np.flipud(image), scale)


def save_pfm_flipped(filename, image_bottom_up, scale=1):
    """The same file from a map whose rows are already in file order (bottom-up)."""
    with open(filename, "wb") as f:
        f.write(_header(image_bottom_up, scale))
        np.ascontiguousarray(image_bottom_up).tofile(f)
