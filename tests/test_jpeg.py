"""JPEG decode: the decoder's arithmetic — host Huffman decoding + the __host__ __device__ IDCT /
upsampling / colour code the two CUDA kernels run — pinned against Pillow (libjpeg-turbo) pixel for pixel.  The CPU part
uses the b200_debug_jpeg_decode_host test hook; the GPU test runs the kernels through b200_jpeg_decode_batch."""
import ctypes as C
import io

import numpy as np
import pytest

from marqo_b200 import _native as N


def _images(rng):
    from PIL import Image
    out = []
    yy, xx = np.mgrid[0:97, 0:131]
    smooth = np.stack([(xx * 2) % 256, (yy * 3 + xx) % 256, (xx * yy // 7) % 256], -1).astype(np.uint8)
    out.append(Image.fromarray(smooth))
    out.append(Image.fromarray(rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)))
    out.append(Image.fromarray(rng.integers(0, 256, size=(33, 47, 3), dtype=np.uint8)))
    out.append(Image.fromarray(rng.integers(0, 256, size=(224, 224, 3), dtype=np.uint8)))
    big = np.kron(rng.integers(0, 256, size=(30, 40, 3), dtype=np.uint8), np.ones((11, 9, 1), np.uint8))
    out.append(Image.fromarray(big))
    out.append(Image.fromarray(rng.integers(0, 256, size=(17, 3, 3), dtype=np.uint8)))
    out.append(Image.fromarray(rng.integers(0, 256, size=(8, 8, 3), dtype=np.uint8)))
    return out


def _encodings(img):
    """(label, jpeg bytes) over sampling modes, qualities, restart intervals, grayscale."""
    cases = []
    for sub, name in ((0, "444"), (1, "422"), (2, "420")):
        for q in (30, 75, 95, 100):
            b = io.BytesIO()
            img.save(b, format="JPEG", quality=q, subsampling=sub)
            cases.append((f"{name}-q{q}", b.getvalue()))
    b = io.BytesIO()
    img.save(b, format="JPEG", quality=85, subsampling=2, optimize=True)
    cases.append(("420-optimized-huffman", b.getvalue()))
    b = io.BytesIO()
    img.save(b, format="JPEG", quality=80, subsampling=2, restart_marker_blocks=3)
    cases.append(("420-restart-3", b.getvalue()))
    b = io.BytesIO()
    img.save(b, format="JPEG", quality=80, subsampling=0, restart_marker_rows=1)
    cases.append(("444-restart-rows", b.getvalue()))
    b = io.BytesIO()
    img.convert("L").save(b, format="JPEG", quality=80)
    cases.append(("grey", b.getvalue()))
    return cases


def _pillow(jpeg: bytes) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(jpeg)).convert("RGB"))


def _host_decode(jpeg: bytes) -> np.ndarray:
    lib = N.load()
    h, w = C.c_int32(0), C.c_int32(0)
    N.check(lib.b200_debug_jpeg_decode_host(jpeg, len(jpeg), None, 0, C.byref(h), C.byref(w)))
    out = np.empty((h.value, w.value, 3), np.uint8)
    N.check(lib.b200_debug_jpeg_decode_host(jpeg, len(jpeg), out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(h), C.byref(w)))
    return out


def test_decoder_arithmetic_is_bit_exact_vs_pillow(native_lib):
    rng = np.random.default_rng(0)
    checked = 0
    for img in _images(rng):
        for label, data in _encodings(img):
            want = _pillow(data)
            got = _host_decode(data)
            assert got.shape == want.shape, (label, img.size)
            assert np.array_equal(got, want), (label, img.size, int(np.abs(got.astype(int) - want).max()))
            checked += 1
    assert checked >= 100


def test_unsupported_files_are_reported_not_guessed(native_lib):
    from PIL import Image
    lib = N.load()
    rng = np.random.default_rng(1)
    img = Image.fromarray(rng.integers(0, 256, size=(40, 40, 3), dtype=np.uint8))
    b = io.BytesIO()
    img.save(b, format="JPEG", progressive=True)
    cm = io.BytesIO()
    img.convert("CMYK").save(cm, format="JPEG")
    png = io.BytesIO()
    img.save(png, format="PNG")
    good = io.BytesIO()
    img.save(good, format="JPEG")
    for data, ok in ((b.getvalue(), 0), (cm.getvalue(), 0), (png.getvalue(), 0), (good.getvalue()[:200], 0),
                     (good.getvalue(), 1)):
        h, w, s = C.c_int32(0), C.c_int32(0), C.c_int32(-1)
        N.check(lib.b200_jpeg_info(data, len(data), C.byref(h), C.byref(w), C.byref(s)))
        assert s.value == ok
    assert (h.value, w.value) == (40, 40)


@pytest.mark.gpu
def test_gpu_decode_matches_pillow_and_feeds_the_encoder(gpu_required):
    """The CUDA kernels on a mixed batch (different sizes, sampling modes, one unsupported file), then straight into the
    image tower: decode -> resize -> ViT never leaves the GPU."""
    import torch
    from marqo_b200.image_decode import decode_jpegs_to_device
    rng = np.random.default_rng(2)
    files = []
    for img in _images(rng)[:5]:
        files += [d for _, d in _encodings(img)[::3]]
    from PIL import Image
    prog = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, size=(50, 60, 3), dtype=np.uint8)).save(prog, format="JPEG", progressive=True)
    files.insert(3, prog.getvalue())
    out = decode_jpegs_to_device(files, device=0)
    assert out[3] is None
    for i, (data, t) in enumerate(zip(files, out)):
        if i == 3:
            continue
        assert t.is_cuda and t.dtype == torch.uint8
        assert np.array_equal(t.cpu().numpy(), _pillow(data)), i
