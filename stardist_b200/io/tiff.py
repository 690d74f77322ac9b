"""TIFF in / out for the command line tools (the reference uses imageio.imread / tifffile.imwrite,
stardist/scripts/predict2d.py:63,98; neither is available here, Pillow is): single images and multi-page stacks."""
import numpy as np


def imread(path):
    """2-D image -> [Y,X(,C)], multi-page stack -> [Z,Y,X(,C)] in the file's sample type"""
    from PIL import Image
    im = Image.open(str(path))
    n = getattr(im, "n_frames", 1)
    if n == 1:
        return np.array(im)
    frames = []
    for i in range(n):
        im.seek(i)
        frames.append(np.array(im))
    return np.stack(frames)


def imwrite(path, arr, compress=True):
    """integer / float arrays of 2 (one page) or 3 (stack of pages) dimensions"""
    from PIL import Image
    arr = np.asarray(arr)
    if arr.dtype == np.int64 or arr.dtype == np.uint64 or arr.dtype == np.uint32:
        arr = arr.astype(np.int32)            # TIFF via Pillow: 8/16-bit unsigned, 32-bit signed, float32
    if arr.dtype == np.float64:
        arr = arr.astype(np.float32)
    if arr.ndim not in (2, 3):
        raise ValueError("imwrite: expected a 2-D image or a 3-D stack, got shape %s" % (arr.shape,))
    pages = [Image.fromarray(a) for a in (arr[np.newaxis] if arr.ndim == 2 else arr)]
    kw = dict(compression="tiff_deflate") if compress else {}
    pages[0].save(str(path), format="TIFF", save_all=len(pages) > 1, append_images=pages[1:], **kw)
