"""GPU image decode for the ingest path.

The reference decodes on its download threads: `Image.open(BytesIO(content))`
(src/marqo/core/inference/image_download.py:146-152) is lazy, the pixels are produced when the model's preprocessor runs
on that thread (src/marqo/tensor_search/add_docs.py:129-134).  Here the preprocessor (loaders._PreprocessToU8) only pulls
the still-encoded bytes out of the lazy PIL image and returns an `EncodedImage`; `encode_image` then decodes the whole
batch with b200_jpeg_decode_batch — Huffman decoding on host threads, IDCT / chroma upsampling / colour conversion on the
GPU, bit-exact with Pillow — and the decoded uint8 HWC tensors go straight into the resize + patch-embed kernels without
visiting the host.  Files outside the decoder's subset (progressive, CMYK, PNG, ...) are decoded by Pillow as before.
"""
from __future__ import annotations

import ctypes as C
import io
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N


class EncodedImage:
    """A still-encoded image travelling through Marqo's media repository.  `.to(device)` is what add_docs.py:130-134
    calls on the preprocessor's result; there is nothing to move yet."""

    __slots__ = ("data", "format")

    def __init__(self, data: bytes, fmt: Optional[str] = None):
        self.data = bytes(data)
        self.format = fmt

    def to(self, device=None):
        return self

    def __len__(self):
        return len(self.data)


def encoded_bytes_of(pil_image) -> Optional[bytes]:
    """The original file bytes of a PIL image that has not been decoded yet (Image.open is lazy), or None."""
    if getattr(pil_image, "format", None) != "JPEG":
        return None
    # already loaded: the bytes may be gone / the image edited.  Pillow >= 11 keeps the core image in `_im` (and `im` is
    # an asserting property); older releases (the reference pins 10.4) use a plain `im` attribute.
    try:
        core = pil_image.__dict__["_im"] if "_im" in pil_image.__dict__ else pil_image.__dict__.get("im")
    except AttributeError:
        core = None
    if core is not None:
        return None
    fp = getattr(pil_image, "fp", None)
    if fp is None or not hasattr(fp, "seek"):
        return None
    try:
        pos = fp.tell()
        fp.seek(0)
        data = fp.read()
        fp.seek(pos)
    except (OSError, ValueError):
        return None
    return data if data[:2] == b"\xff\xd8" else None


def pillow_decode(data: bytes) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"), dtype=np.uint8)


def decode_jpegs_to_device(files: Sequence[bytes], device: int = 0) -> List[Optional["torch.Tensor"]]:
    """Decode a batch of JPEG files on GPU `device`.  -> one uint8 [H, W, 3] CUDA tensor per file, or None where the
    file is outside the decoder's subset (the caller decodes those with Pillow)."""
    import torch
    lib = N.load()
    n = len(files)
    if n == 0:
        return []
    bufs = [C.create_string_buffer(f, len(f)) for f in files]
    ptrs = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in bufs])
    sizes = (C.c_size_t * n)(*[len(f) for f in files])
    heights = (C.c_int32 * n)()
    widths = (C.c_int32 * n)()
    status = (C.c_int32 * n)()
    ok, outs = [], [None] * n
    for i in range(n):     # size pass: headers only matter, but support is only certain after a full parse
        h, w, s = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        N.check(lib.b200_jpeg_info(C.cast(bufs[i], C.c_void_p), len(files[i]), C.byref(h), C.byref(w), C.byref(s)))
        if s.value:
            outs[i] = torch.empty((h.value, w.value, 3), dtype=torch.uint8, device=torch.device("cuda", device))
            ok.append(i)
    if not ok:
        return outs
    torch.cuda.synchronize(device)
    d_out = (C.c_void_p * n)(*[(outs[i].data_ptr() if outs[i] is not None else None) for i in range(n)])
    N.check(lib.b200_jpeg_decode_batch(int(device), ptrs, sizes, n, d_out, heights, widths, status))
    for i in range(n):
        if outs[i] is not None and status[i] != N.OK:
            outs[i] = None
    return outs


def decode_images_to_device(items: Sequence, device: int = 0) -> List["torch.Tensor"]:
    """EncodedImage / bytes -> uint8 HWC CUDA tensors: JPEGs on the GPU, everything else through Pillow + one H2D."""
    import torch
    datas = [it.data if isinstance(it, EncodedImage) else bytes(it) for it in items]
    outs = decode_jpegs_to_device(datas, device)
    dev = torch.device("cuda", device)
    for i, t in enumerate(outs):
        if t is None:
            outs[i] = torch.from_numpy(pillow_decode(datas[i])).to(dev)
    return outs
