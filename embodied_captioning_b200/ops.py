"""Torch-tensor front-ends of the exported kernels (thin: argument checking + ctypes call).

PyTorch only owns the device memory and the stream here; every operation runs in `libec_b200.so`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as nat

EPI_BF16, EPI_BF16_GELU, EPI_F32, EPI_F32_RESID = 0, 1, 2, 3
MODE_FULL_MEAN_1MINUS, MODE_PAIRS_MEAN_COS = 0, 1
DEDUP = 4            # EC_DISAGREEMENT_DEDUP: OR into the mode (projection_utils.py:574-578 `unique(dim=0)`)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_dev(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise nat.NativeError("expected CUDA tensors (there is no CPU path)")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, epilogue: int = EPI_BF16,
         residual: Optional[torch.Tensor] = None, simt_reference: bool = False) -> torch.Tensor:
    """out[M,N] = a[M,K] @ w[N,K]^T (+bias) with the fused epilogue; a, w bf16 row-major."""
    _check_dev(a, w, bias, residual)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=torch.bfloat16 if epilogue in (EPI_BF16, EPI_BF16_GELU)
                      else torch.float32)
    with torch.cuda.device(a.device):
        nat.check(nat.lib().ec_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, nat.ptr(bias),
                                         nat.ptr(residual), residual.stride(0) if residual is not None else 0,
                                         out.data_ptr(), out.stride(0), epilogue, int(simt_reference), _stream(a)),
                  "ec_gemm_bf16")
    return out


def gemm_splitk(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], residual: torch.Tensor, split_k: int = 0,
                counters: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] f32 = residual + a @ w^T (+ bias) with deterministic split-K (2..4 slices per tile; 0 = choose)."""
    _check_dev(a, w, bias, residual)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    ws = torch.empty((3 * M * N,), device=a.device, dtype=torch.float32)
    if counters is None:
        counters = torch.zeros(((M + 127) // 128 * ((N + 127) // 128),), device=a.device, dtype=torch.int32)
    with torch.cuda.device(a.device):
        nat.check(nat.lib().ec_gemm_bf16_splitk(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K,
                                                nat.ptr(bias), residual.data_ptr(), residual.stride(0), out.data_ptr(),
                                                out.stride(0), split_k, ws.data_ptr(), counters.data_ptr(), _stream(a)),
                  "ec_gemm_bf16_splitk")
    return out


def gemm_grouped(a: torch.Tensor, w: torch.Tensor, group_cols: int, bias: Optional[torch.Tensor] = None,
                 epilogue: int = EPI_BF16, simt_reference: bool = False) -> torch.Tensor:
    """out[m, n] = sum_k a[m, (n // group_cols) * K + k] * w[n, k] (+ bias): a [M, groups*K], w [N, K] bf16."""
    _check_dev(a, w, bias)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    M, N, K = a.shape[0], w.shape[0], w.shape[1]
    assert N % group_cols == 0 and a.shape[1] == (N // group_cols) * K
    out = torch.empty((M, N), device=a.device, dtype=torch.bfloat16 if epilogue == EPI_BF16 else torch.float32)
    with torch.cuda.device(a.device):
        nat.check(nat.lib().ec_gemm_bf16_grouped(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K,
                                                 group_cols, nat.ptr(bias), out.data_ptr(), out.stride(0), epilogue,
                                                 int(simt_reference), _stream(a)), "ec_gemm_bf16_grouped")
    return out


def pack_image_tokens(x: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """x f32 [B, n_kv, T] -> packed bf16 buffer (32-row chunks, rows of T + 8 elements, zero padded) that
    `cross_attention_absorbed` streams; normalize = LayerNorm without affine (eps 1e-5)."""
    _check_dev(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    B, n_kv, T = x.shape
    lib = nat.lib()
    out = torch.empty((lib.ec_packed_image_tokens_bytes(B, n_kv, T) // 2,), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        nat.check(lib.ec_pack_image_tokens(x.data_ptr(), out.data_ptr(), B, n_kv, T, int(normalize), _stream(x)),
                  "ec_pack_image_tokens")
    return out


def cross_attention_absorbed(qp: torch.Tensor, xn_packed: torch.Tensor, n_kv: int,
                             finished: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                             reverse: bool = False) -> torch.Tensor:
    """qp bf16 [B, H, T], xn_packed = pack_image_tokens(...) -> ctx bf16 [B, H, T]: softmax(qp xn^T) xn per crop."""
    _check_dev(qp, xn_packed, finished)
    assert qp.dtype == torch.bfloat16 and xn_packed.dtype == torch.bfloat16 and qp.is_contiguous()
    B, H, T = qp.shape
    assert xn_packed.numel() * 2 == nat.lib().ec_packed_image_tokens_bytes(B, n_kv, T)
    ctx = torch.zeros_like(qp) if out is None else out
    with torch.cuda.device(qp.device):
        nat.check(nat.lib().ec_cross_attention_absorbed(qp.data_ptr(), xn_packed.data_ptr(), ctx.data_ptr(), B, H, T,
                                                        n_kv, nat.ptr(finished), int(reverse), _stream(qp)),
                  "ec_cross_attention_absorbed")
    return ctx


def gemm_argmax(a: torch.Tensor, w: torch.Tensor, want_logits: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    _check_dev(a, w)
    M, K = a.shape
    N = w.shape[0]
    logits = torch.empty((M, N), device=a.device, dtype=torch.float32) if want_logits else None
    arg = torch.empty((M,), device=a.device, dtype=torch.int32)
    parts = (N + 255) // 256 * 2
    ws = torch.empty((M * parts * 8 + 512,), device=a.device, dtype=torch.uint8)
    with torch.cuda.device(a.device):
        nat.check(nat.lib().ec_gemm_argmax(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K,
                                           nat.ptr(logits), arg.data_ptr(), ws.data_ptr(), ws.numel(), _stream(a)),
                  "ec_gemm_argmax")
    return arg, logits


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_bf16: bool = True) -> torch.Tensor:
    _check_dev(x, gamma, beta)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, width = x.shape
    y = torch.empty((rows, width), device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    with torch.cuda.device(x.device):
        nat.check(nat.lib().ec_layernorm(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                         y.stride(0), int(out_bf16), rows, width, _stream(x)), "ec_layernorm")
    return y


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """q [B,Sq,H*D] (or [Sq,H*D] shared by the batch), k/v [B,Skv,H*D] bf16 -> [B,Sq,H*D] bf16."""
    _check_dev(q, k, v)
    B, Skv, C = k.shape
    D = C // heads
    if q.dim() == 2:
        q_bs, Sq = 0, q.shape[0]
        q_rs = q.stride(0)
    else:
        q_bs, Sq, q_rs = q.stride(0), q.shape[1], q.stride(1)
    out = torch.empty((B, Sq, C), device=k.device, dtype=torch.bfloat16)
    with torch.cuda.device(k.device):
        nat.check(nat.lib().ec_attention(q.data_ptr(), q_bs, q_rs, k.data_ptr(), k.stride(0), k.stride(1), v.data_ptr(),
                                         v.stride(0), v.stride(1), out.data_ptr(), out.stride(0), out.stride(1), B,
                                         heads, Sq, Skv, D, _stream(k)), "ec_attention")
    return out


def attention_fused_qkv(qkv: torch.Tensor, heads: int, impl: int = 1) -> torch.Tensor:
    """qkv bf16 [B, S, 3*H*D] (q | k | v column blocks) -> [B, S, H*D]; impl 1 = tcgen05/TMEM kernel, 0 = mma.sync."""
    _check_dev(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.dim() == 3 and qkv.is_contiguous()
    B, S, C3 = qkv.shape
    C = C3 // 3
    out = torch.empty((B, S, C), device=qkv.device, dtype=torch.bfloat16)
    with torch.cuda.device(qkv.device):
        nat.check(nat.lib().ec_attention_fused_qkv(qkv.data_ptr(), C3, 0, C, 2 * C, out.data_ptr(), S * C, C, B, heads, S,
                                                   C // heads, impl, _stream(qkv)), "ec_attention_fused_qkv")
    return out


def preprocess_crops(frames: torch.Tensor, rects: torch.Tensor, frame_idx: torch.Tensor,
                     flags: Optional[torch.Tensor] = None, size: int = 224, want_f32: bool = False,
                     want_bf16: bool = True):
    """frames uint8 [F,H,W,C] on device; rects int32 [N,4]; -> (bf16 [N,3,S,S] | None, f32 | None, status int32 [N])."""
    _check_dev(frames, rects, frame_idx, flags)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_contiguous()
    assert rects.dtype == torch.int32 and rects.is_contiguous() and frame_idx.dtype == torch.int32
    N = rects.shape[0]
    F_, H, W, Cc = frames.shape
    dev = frames.device
    out_bf16 = torch.empty((N, 3, size, size), device=dev, dtype=torch.bfloat16) if want_bf16 else None
    out_f32 = torch.empty((N, 3, size, size), device=dev, dtype=torch.float32) if want_f32 else None
    status = torch.zeros((max(N, 1),), device=dev, dtype=torch.int32)
    lib = nat.lib()
    with torch.cuda.device(dev):
        for s in range(0, N, 32768):      # grid.y limit
            e = min(N, s + 32768)
            nat.check(lib.ec_preprocess_crops(
                frames.data_ptr(), F_, H, W, Cc, rects[s:e].data_ptr(), frame_idx[s:e].data_ptr(),
                None if flags is None else flags[s:e].data_ptr(), e - s, size,
                None if out_bf16 is None else out_bf16[s:e].data_ptr(),
                None if out_f32 is None else out_f32[s:e].data_ptr(), status[s:e].data_ptr(), _stream(frames)),
                "ec_preprocess_crops")
    return out_bf16, out_f32, status[:N]


def instance_disagreement(emb: torch.Tensor, keys: torch.Tensor, mode: int = MODE_FULL_MEAN_1MINUS, dedup: bool = False):
    """emb f32 [N,D], keys int64 [N] on device -> (keys[M] int64, score[M] f32, count[M] int32), key-sorted.
    dedup=True removes exact-duplicate views of an instance first (count = distinct views)."""
    mode = int(mode) | (DEDUP if dedup else 0)
    _check_dev(emb, keys)
    assert emb.dtype == torch.float32 and emb.dim() == 2 and emb.stride(1) == 1
    assert keys.dtype == torch.int64 and keys.is_contiguous()
    N, D = emb.shape
    dev = emb.device
    ok = torch.empty((max(N, 1),), device=dev, dtype=torch.int64)
    sc = torch.empty((max(N, 1),), device=dev, dtype=torch.float32)
    ct = torch.empty((max(N, 1),), device=dev, dtype=torch.int32)
    m = torch.zeros((1,), device=dev, dtype=torch.int32)
    lib = nat.lib()
    wsb = lib.ec_disagreement_workspace_bytes(N)
    ws = torch.empty((max(wsb, 16),), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        nat.check(lib.ec_instance_disagreement(emb.data_ptr(), emb.stride(0) if N > 1 else D, keys.data_ptr(), N, D, mode,
                                               ok.data_ptr(),
                                               sc.data_ptr(), ct.data_ptr(), m.data_ptr(), ws.data_ptr(), ws.numel(),
                                               _stream(emb)), "ec_instance_disagreement")
    M = int(m.item())
    return ok[:M], sc[:M], ct[:M]


def check_watchdog(device=None):
    with torch.cuda.device(device):
        nat.check(nat.lib().ec_check_watchdog(torch.cuda.current_stream(device).cuda_stream), "ec_check_watchdog")


def caption_frequency(tokens: torch.Tensor, keys: torch.Tensor):
    """tokens int [N,T], keys int64 [N] on device -> (key[M] int64, first_row[M] int32, count[M] int32): one entry per
    distinct (instance, caption), ordered by (key, first occurrence) -- `compute_captions_frequency`
    (pseudocaptioner.py:156-177) on token rows."""
    _check_dev(tokens, keys)
    tok = tokens.to(torch.int32).contiguous()
    k = keys.to(torch.int64).contiguous()
    N, T = tok.shape
    dev = tok.device
    ok = torch.empty((max(N, 1),), device=dev, dtype=torch.int64)
    of = torch.empty((max(N, 1),), device=dev, dtype=torch.int32)
    oc = torch.empty((max(N, 1),), device=dev, dtype=torch.int32)
    m = torch.zeros((1,), device=dev, dtype=torch.int32)
    lib = nat.lib()
    ws = torch.empty((max(lib.ec_caption_frequency_workspace_bytes(N), 16),), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        nat.check(lib.ec_caption_frequency(tok.data_ptr() if N else None, tok.stride(0) if N else T, T,
                                           k.data_ptr() if N else None, N, ok.data_ptr(), of.data_ptr(), oc.data_ptr(),
                                           m.data_ptr(), ws.data_ptr(), ws.numel(), _stream(tok)), "ec_caption_frequency")
    M = int(m.item())
    return ok[:M], of[:M], oc[:M]
