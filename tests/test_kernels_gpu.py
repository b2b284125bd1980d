"""Kernel-level numerics tests: each CUDA kernel of the encoder vs a plain PyTorch fp32 reference of the same op
(inputs pre-rounded to bf16 where the kernel consumes bf16, so the comparison isolates the kernel's arithmetic)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64), (128, 64, 64), (200, 96, 128), (50, 512, 768), (1000, 768, 768), (257 * 3, 3072, 1024),
    (4096, 1024, 4096), (392, 768, 3072), (12800, 2304, 768), (300, 128, 640),
])
def test_gemm_matches_torch(gpu_required, M, N, K):
    from marqo_b200.engine import debug_gemm
    g = torch.Generator().manual_seed(M + N + K)
    A = _bf16(torch.randn(M, K, generator=g))
    W = _bf16(torch.randn(N, K, generator=g) / math.sqrt(K))
    bias = torch.randn(N, generator=g)
    ref = A @ W.t() + bias
    got = torch.from_numpy(debug_gemm(A.numpy(), W.numpy(), bias.numpy()))
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)   # fp32 accumulate, different summation order


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_epilogues(gpu_required, act):
    from marqo_b200.engine import debug_gemm
    g = torch.Generator().manual_seed(act)
    M, N, K = 333, 1024, 256
    A = _bf16(torch.randn(M, K, generator=g))
    W = _bf16(torch.randn(N, K, generator=g) / math.sqrt(K))
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    z = A @ W.t() + bias
    a = torch.nn.functional.gelu(z) if act == 1 else z * torch.sigmoid(1.702 * z)
    got = torch.from_numpy(debug_gemm(A.numpy(), W.numpy(), bias.numpy(), res.numpy(), act=act))
    torch.testing.assert_close(got, a + res, rtol=2e-4, atol=3e-4)
    got_b = torch.from_numpy(debug_gemm(A.numpy(), W.numpy(), bias.numpy(), None, act=act, out_bf16=True))
    torch.testing.assert_close(got_b, a, rtol=1e-2, atol=1e-2)     # bf16 output rounding


@pytest.mark.parametrize("M,N,K,in_place", [
    (257 * 5, 1024, 1024, False),    # ViT-L out_proj shape: 4 N tiles of 256, 8 writers per 32-row strip
    (1000, 768, 3072, False),        # ViT-B fc2: 3 N tiles; M % 32 != 0 (a short last strip)
    (77 * 3, 512, 512, False),       # CLIP text width 512
    (300, 128, 256, True),           # BN = 128 tile, in-place fp32 (BERT post-LN)
    (128 * 150 + 17, 1024, 256, True),   # more tiles than CTA pairs: every CTA finishes strips of several row bands
    (40, 384, 128, False),           # N = 384 on the 256-wide tile: the last tile's second half has no columns
])
def test_gemm_fused_layernorm(gpu_required, M, N, K, in_place):
    """LayerNorm inside the residual GEMM's epilogue (last writer of a 32-row strip normalises it) vs torch."""
    from marqo_b200.engine import debug_gemm_ln
    g = torch.Generator().manual_seed(M + N + K)
    A = _bf16(torch.randn(M, K, generator=g))
    W = _bf16(torch.randn(N, K, generator=g) / math.sqrt(K))
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    gamma = 1.0 + 0.1 * torch.randn(N, generator=g)
    beta = 0.1 * torch.randn(N, generator=g)
    eps = 1e-12 if in_place else 1e-5
    x_ref = (A.double() @ W.double().t() + bias.double() + res.double())
    ln_ref = torch.nn.functional.layer_norm(x_ref, (N,), gamma.double(), beta.double(), eps)
    # repeats = 3: the strip counters must return to zero after every launch (not in place: same result each time)
    x, ln = debug_gemm_ln(A.numpy(), W.numpy(), bias.numpy(), res.numpy(), gamma.numpy(), beta.numpy(), eps,
                          in_place=in_place, repeats=1 if in_place else 3)
    x, ln = torch.from_numpy(x).double(), torch.from_numpy(ln).double()
    torch.testing.assert_close(ln, ln_ref, rtol=1e-2, atol=1e-2)            # bf16 output rounding
    torch.testing.assert_close(x, ln_ref if in_place else x_ref, rtol=3e-4, atol=3e-4)


@pytest.mark.parametrize("n,S,patch,N", [
    (3, 224, 14, 1024),    # ViT-L-14: 42-byte pixel rows, one 64-slot k-block each, 3 of its 4 UMMA_K steps issued
    (5, 224, 32, 768),     # ViT-B-32: 96-byte pixel rows = 2 k-blocks, a warp's 32 patches straddle images (49 per image)
    (2, 224, 16, 128),     # ViT-B-16 grid with the BN = 128 tile
    (300, 224, 32, 128),   # more tiles than CTA pairs: the smem ring and the strip buffers wrap
    (1, 112, 8, 256),      # small image: 336-byte rows
])
def test_patch_embed_gather_matches_conv(gpu_required, n, S, patch, N):
    """uint8 HWC -> ToTensor -> Normalize -> conv1 fused into the GEMM's operand load.  Reference: the
    torchvision formula (u8/255 - mean)/std in fp32, rounded to bf16 like the kernel's A operand, conv2d in fp32."""
    from marqo_b200.engine import debug_patch_embed
    g = torch.Generator().manual_seed(n * 31 + patch)
    img = torch.randint(0, 256, (n, S, S, 3), generator=g, dtype=torch.uint8)
    K = 3 * patch * patch
    w = _bf16(torch.randn(N, 3, patch, patch, generator=g) / math.sqrt(K))
    G = (S // patch) ** 2
    pos = torch.randn(G + 1, N, generator=g)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073])
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711])
    x = (img.permute(0, 3, 1, 2).float() / 255.0 - mean[None, :, None, None]) / std[None, :, None, None]
    ref = torch.nn.functional.conv2d(_bf16(x).double(), w.double(), stride=patch)       # [n, N, g, g]
    ref = ref.flatten(2).transpose(1, 2) + pos[None, 1:, :].double()                      # [n, G, N]
    got = torch.from_numpy(debug_patch_embed(img.numpy(), patch, w.numpy(), mean.numpy(), std.numpy(), pos.numpy()))
    got = got.view(n, G + 1, N)
    assert float(got[:, 0].abs().max()) == 0.0                                            # class rows are not this kernel's
    # the kernel normalises with one fma (u * 1/(255 std) - mean/std): a few values land on the other side of a bf16
    # rounding boundary (2^-9 relative) -> compare at bf16-product accuracy, and against the im2col path the same way
    torch.testing.assert_close(got[:, 1:].double(), ref, rtol=0, atol=2e-2)
    assert float((got[:, 1:].double() - ref).abs().mean()) < 1e-3
    old = torch.from_numpy(debug_patch_embed(img.numpy(), patch, w.numpy(), mean.numpy(), std.numpy(), pos.numpy(),
                                             use_gather=False)).view(n, G + 1, N)
    torch.testing.assert_close(got, old, rtol=0, atol=2e-2)


@pytest.mark.parametrize("B,S,H,mask", [
    (2, 50, 12, 0), (3, 257, 4, 0), (2, 77, 8, 1), (4, 128, 12, 2), (2, 512, 2, 2), (1, 1, 2, 0), (2, 64, 2, 1), (1, 65, 2, 1),
    (2, 129, 2, 0), (2, 136, 2, 2), (2, 137, 2, 0), (3, 385, 2, 2), (2, 129, 2, 1), (2, 260, 2, 1), (1, 300, 2, 1), (2, 256, 4, 0),
    (1, 1025, 2, 0),
    # more work items than the persistent grid (2 x 148 CTAs): every CTA loops over several (batch, head, query block)
    # items, so barrier phases, the K/V ring and the remainder-key staging wrap around
    (40, 257, 8, 0), (32, 385, 4, 2), (160, 129, 2, 1), (80, 128, 4, 2), (12, 512, 8, 2), (100, 130, 3, 0),
    # the one-shot kernel (129 <= S <= 257, attention_os.cu): key-length masks shorter than one / two tiles, the 257th
    # token with a key-length mask, more (batch, head) units than CTAs (Q ring, K / V hand-back and TMEM reuse wrap)
    (5, 200, 2, 2), (6, 257, 2, 2), (3, 256, 2, 2), (90, 257, 4, 0), (170, 197, 2, 2), (2, 255, 3, 0),
])
def test_attention_matches_torch(gpu_required, B, S, H, mask):
    from marqo_b200.engine import debug_attention
    g = torch.Generator().manual_seed(B * 1000 + S)
    W = H * 64
    qkv = _bf16(torch.randn(B * S, 3 * W, generator=g))
    kv_len = None
    if mask == 2:
        kv_len = torch.randint(1, S + 1, (B,), generator=g).to(torch.int32)
        kv_len[0] = S
    q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = (q @ k.transpose(-1, -2)) / 8.0
    if mask == 1:
        att = att + torch.full((S, S), float("-inf")).triu_(1)
    if mask == 2:
        keep = torch.arange(S)[None, :] < kv_len[:, None]
        att = att.masked_fill(~keep[:, None, None, :], float("-inf"))
    ref = (att.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * S, W)
    got = torch.from_numpy(debug_attention(qkv.numpy(), B, S, W, H, mask, None if kv_len is None else kv_len.numpy()))
    torch.testing.assert_close(got, ref, rtol=2e-2, atol=2e-2)     # P and the output are rounded to bf16
    assert (got - ref).abs().mean() < 3e-3


@pytest.mark.parametrize("B,S,H,mask", [(3, 257, 2, 0), (4, 200, 2, 2), (3, 512, 2, 2), (4, 77, 2, 1)])
def test_attention_peaked_scores(gpu_required, B, S, H, mask):
    """Scores with a spread of +-40 (one key dominates most rows): the exponent reference must be the row's true maximum."""
    from marqo_b200.engine import debug_attention
    g = torch.Generator().manual_seed(S)
    W = H * 64
    qkv = torch.randn(B * S, 3 * W, generator=g)
    qkv[:, : 2 * W] *= 3.0                                          # q and k: score std 9, extremes beyond 40
    qkv = _bf16(qkv)
    kv_len = None
    if mask == 2:
        kv_len = torch.randint(1, S + 1, (B,), generator=g).to(torch.int32)
        kv_len[0] = S
    q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = (q.double() @ k.double().transpose(-1, -2)) / 8.0
    if mask == 1:
        att = att + torch.full((S, S), float("-inf"), dtype=torch.float64).triu_(1)
    if mask == 2:
        keep = torch.arange(S)[None, :] < kv_len[:, None]
        att = att.masked_fill(~keep[:, None, None, :], float("-inf"))
    ref = (att.softmax(-1) @ v.double()).permute(0, 2, 1, 3).reshape(B * S, W).float()
    got = torch.from_numpy(debug_attention(qkv.numpy(), B, S, W, H, mask, None if kv_len is None else kv_len.numpy()))
    assert torch.isfinite(got).all()
    torch.testing.assert_close(got, ref, rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("rows,w,eps", [(5, 128, 1e-5), (77, 512, 1e-5), (1000, 768, 1e-12), (33, 1024, 1e-5)])
def test_layernorm_matches_torch(gpu_required, rows, w, eps):
    from marqo_b200.engine import debug_layernorm
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, w, generator=g) * 3 + 1
    gamma, beta = torch.randn(w, generator=g), torch.randn(w, generator=g)
    ref = torch.nn.functional.layer_norm(x, (w,), gamma, beta, eps)
    got = torch.from_numpy(debug_layernorm(x.numpy(), gamma.numpy(), beta.numpy(), eps))
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("h,w", [(480, 640), (640, 480), (224, 224), (300, 224), (256, 256), (1000, 750), (225, 400), (100, 150)])
def test_resize_matches_pillow_bit_exact(gpu_required, h, w):
    """Resize(224, BICUBIC) + CenterCrop(224) on PIL images (clip_utils.py:48-67) — Pillow is the third-party
    implementation the reference runs; the CUDA kernel restates its fixed-point two-pass resampler bit for bit."""
    from PIL import Image
    from torchvision.transforms import CenterCrop, InterpolationMode, Resize
    from marqo_b200.engine import debug_resize
    rng = np.random.default_rng(h * 7 + w)
    imgs = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
    imgs[1] = (np.linspace(0, 255, w)[None, :, None] * np.ones((h, 1, 3))).astype(np.uint8)   # smooth gradient
    tf = [Resize(224, interpolation=InterpolationMode.BICUBIC), CenterCrop(224)]
    ref = []
    for a in imgs:
        im = Image.fromarray(a)
        for t in tf:
            im = t(im)
        ref.append(np.asarray(im.convert("RGB")))
    got = debug_resize(imgs, 224)
    np.testing.assert_array_equal(got, np.stack(ref))
