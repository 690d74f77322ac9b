"""ImageJ polygon ROI export (stardist/utils.py:196-268: polyroi_bytearray, export_imagej_rois).

Binary layout of one .roi record as read by ij/io/RoiDecoder.java (all fields big endian):
  0 "Iout" | 4 int16 version (227) | 6 int16 type (0 = polygon) | 8 top | 10 left | 12 bottom | 14 right (int16 bbox)
  16 uint16 n | 50 int16 options (128 = sub-pixel resolution) | 56 int32 position | 64 int16 x[n] - left, then y[n] - top
  then, with sub-pixel resolution, float32 x[n], float32 y[n] (absolute).  ImageJ's pixel centres sit at +0.5."""
import struct
from pathlib import Path
from zipfile import ZipFile, ZIP_DEFLATED
import numpy as np

_HEADER = 64


def polyroi_bytearray(x, y, pos=None, subpixel=True):
    xf = np.asarray(x, dtype=np.float64).ravel() + 0.5
    yf = np.asarray(y, dtype=np.float64).ravel() + 0.5
    if len(xf) != len(yf):
        raise ValueError("x and y must have the same length")
    xi, yi = np.round(xf), np.round(yf)
    n = len(xi)
    top, left, bottom, right = yi.min(), xi.min(), yi.max(), xi.max()
    buf = bytearray(_HEADER + 4 * n + (8 * n if subpixel else 0))
    struct.pack_into(">4shhhhhhH", buf, 0, b"Iout", 227, 0, int(top), int(left), int(bottom), int(right), n)
    if subpixel:
        struct.pack_into(">h", buf, 50, 128)
    if pos is not None:
        struct.pack_into(">i", buf, 56, int(pos))
    struct.pack_into(">%dh" % n, buf, _HEADER, *[int(v - left) for v in xi])
    struct.pack_into(">%dh" % n, buf, _HEADER + 2 * n, *[int(v - top) for v in yi])
    if subpixel:
        struct.pack_into(">%df" % n, buf, _HEADER + 4 * n, *xf.tolist())
        struct.pack_into(">%df" % n, buf, _HEADER + 8 * n, *yf.tolist())
    return buf


def export_imagej_rois(fname, polygons, set_position=True, subpixel=True, compression=ZIP_DEFLATED):
    """polygons: array [n,2,R] of (y,x) coordinates (res_dict['coord']) or a sequence of such arrays (one per frame)
    -> `<fname>.zip` with one `PPP_III.roi` entry per polygon"""
    if isinstance(polygons, np.ndarray):
        polygons = (polygons,)
    fname = Path(fname)
    if fname.suffix == ".zip":
        fname = fname.with_suffix("")
    with ZipFile(str(fname) + ".zip", mode="w", compression=compression) as z:
        for pos, group in enumerate(polygons, start=1):
            for i, poly in enumerate(group, start=1):
                z.writestr("%03d_%03d.roi" % (pos, i), bytes(polyroi_bytearray(poly[1], poly[0], pos=(pos if set_position else None), subpixel=subpixel)))
