"""Host logic of the DAMSM drop-ins: autograd plumbing, sharding by caption rows and the collectives.

All arithmetic is done by the engine (``engine.CudaEngine`` -> libdamsm_b200.so).  What lives here is
the reference's *interface* behaviour (DMGAN+CLIP/code/miscc/losses.py:51-91, 219-272;
GlobalAttention.py:38-160) and the multi-GPU exchange steps of SURVEY.md 8(e):

  forward   all_gather(vhat / sentence codes / class ids)  ->  local (B/N x B) block of logits
            -> row LSE (complete) + column (max, sum-exp) partials -> all_reduce -> both CE losses
  backward  d(words) complete locally;  d(regions) partial for all B images -> reduce_scatter.
"""
from __future__ import annotations

import math

import warnings
import weakref

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import get_engine

DEFAULT_GAMMAS = (4.0, 5.0, 10.0)      # cfg/DAMSM/bird.yml:27-29, coco.yml:27-29, clip_bird_DMGAN.yml:25-28
SENT_FUSED_MAX_PAIRS = 128 * 128       # sent_loss: logits up to this size take the one-launch kernels (sent_fused.cu)


# ----------------------------------------------------------------------------------------------- collectives
def _world(group):
    if group is None:
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _all_gather_rows(x: torch.Tensor, group):
    """Concatenate equal-sized row shards from every rank (rank order)."""
    if group is None:
        return x
    world = dist.get_world_size(group)
    x = x.contiguous()
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    dist.all_gather_into_tensor(out, x, group=group)
    return out


def _reduce_scatter_rows(x: torch.Tensor, group):
    """Sum ``x`` (B_total, ...) over ranks and return this rank's row shard."""
    if group is None:
        return x
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    x = x.contiguous()
    rows = x.shape[0] // world
    if dist.get_backend(group) == "gloo":          # gloo has no reduce_scatter
        dist.all_reduce(x, group=group)
        return x[rank * rows:(rank + 1) * rows].clone()
    out = torch.empty((rows,) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    dist.reduce_scatter_tensor(out, x, group=group)
    return out


def combine_column_lse(col_max: torch.Tensor, col_sum: torch.Tensor, group):
    """Column log-sum-exp of the full (B x B) matrix from per-rank (max, sum exp(x-max)) partials."""
    if group is None:
        return torch.log(col_sum) + col_max
    gmax = col_max.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    scale = torch.where(col_max == gmax, torch.ones_like(col_max), torch.exp(col_max - gmax))
    s = col_sum * scale
    dist.all_reduce(s, group=group)
    return torch.log(s) + gmax


# ----------------------------------------------------------------------------------------------- prologue cache
# project_regions() already produces the normalised copies (fp32, fp16) and norms of its output in the GEMM epilogue.
# They are remembered here, keyed by the output's storage, so that a following words_loss on that very tensor (any view
# with the same layout, unmodified) skips its own l2norm pass over the largest tensor of the loss.
_prologue = {}


def _prologue_drop(key, ref):
    hit = _prologue.get(key)
    if hit is not None and hit[0] is ref:          # not already replaced by a newer tensor at the same address
        del _prologue[key]


def _prologue_put(y, side):
    key = y.data_ptr()
    ref = weakref.ref(y)
    _prologue[key] = (ref, y._version, side)
    weakref.finalize(y, _prologue_drop, key, ref)  # the copies (GBs at large batches) die with the projection output


def _prologue_get(regions3):
    hit = _prologue.get(regions3.data_ptr())
    if hit is None:
        return None
    y = hit[0]()
    if (y is None or regions3._version != hit[1] or regions3.shape != y.shape or regions3.stride() != y.stride()
            or regions3.dtype != y.dtype):
        return None
    return hit[2]


# ----------------------------------------------------------------------------------------------- words loss
class DamsmWordsLoss(torch.autograd.Function):
    """words_loss (losses.py:219-272) for the local caption rows against all images."""

    @staticmethod
    def forward(ctx, regions3, words3, mask_u8, labels, cls_local, gammas, engine, group):
        world, rank = _world(group)
        bl = words3.shape[0]
        if regions3.shape[0] != bl:
            raise ValueError("words and regions must have the same (local) batch size")
        b_total, row_offset = bl * world, rank * bl
        qhat, qhat16, qnorm, qunorm = engine.l2norm_fwd(words3, want_bf16=engine.precision == "bf16", pad8=True)
        side = _prologue_get(regions3)
        if side is not None:                     # regions come straight from project_regions(): already normalised
            vhat_l, vhat16_l, vnorm = side[0], (side[1] if engine.precision == "bf16" else None), side[2]
        else:
            vhat_l, vhat16_l, vnorm, _ = engine.l2norm_fwd(regions3, want_bf16=engine.precision == "bf16")
        # image side: Gram matrices are computed (and packed) for the local images only; what is gathered is exactly
        # what the pair kernels read (tensor-core path: fp16 gx + fp16 vhat; exact path: fp32 gram + fp32 vhat)
        cls_all = _all_gather_rows(cls_local, group) if cls_local is not None else None
        if hasattr(engine, "image_side"):
            colside, vhat = engine.image_side(vhat_l, vhat16_l, lambda x: _all_gather_rows(x, group))
        else:                                    # block-level engines without the fused helper (tests/checker_engine.py)
            gram = _all_gather_rows(engine.gram(vhat_l), group)
            tc = engine.precision == "bf16"
            vhat = vhat_l if (group is None or tc) else _all_gather_rows(vhat_l, group)
            vhat16 = _all_gather_rows(vhat16_l, group) if vhat16_l is not None else None
            colside = engine.pack_columns(gram, vhat, vhat16)
        sim = engine.words_fwd(qhat, qhat16, vhat, colside, qunorm, mask_u8, gammas)
        row_lse, col_max, col_sum = engine.ce_stats(sim, cls_local, cls_all, row_offset)
        col_lse = combine_column_lse(col_max, col_sum, group)
        out2 = engine.ce_losses(sim, row_lse, col_lse, labels, row_offset, b_total)
        if group is not None:
            dist.all_reduce(out2, group=group)
        ctx.engine, ctx.group, ctx.gammas = engine, group, gammas
        ctx.row_offset, ctx.b_total = row_offset, b_total
        ctx.colside, ctx.qhat16 = colside, qhat16
        ctx.save_for_backward(regions3, words3, mask_u8, labels, qhat, vhat, vhat_l, qnorm, qunorm, vnorm,
                              sim, row_lse, col_lse)
        ctx.mark_non_differentiable(sim)
        return out2[0].clone(), out2[1].clone(), sim

    @staticmethod
    def backward(ctx, g0, g1, _gsim):
        (regions3, words3, mask_u8, labels, qhat, vhat, vhat_l, qnorm, qunorm, vnorm,
         sim, row_lse, col_lse) = ctx.saved_tensors
        eng, colside = ctx.engine, ctx.colside
        gscale = torch.stack([g0.reshape(()), g1.reshape(())]).to(torch.float32)
        dqhat, dvhat, hmat, kq = eng.words_bwd(qhat, ctx.qhat16, vhat, colside, qunorm, mask_u8, sim, row_lse, col_lse,
                                               labels, gscale, ctx.row_offset, ctx.b_total, ctx.gammas,
                                               need_dq=ctx.needs_input_grad[1], need_dv=ctx.needs_input_grad[0])
        dregions3 = dwords3 = None
        if ctx.needs_input_grad[0]:
            # partial sums over this rank's captions for ALL images -> owners; the -H vhat term is applied there
            dvhat_l = _reduce_scatter_rows(dvhat, ctx.group)
            hmat_l = _reduce_scatter_rows(hmat, ctx.group)
            dvhat_l = eng.gram_bwd(hmat_l, vhat_l, dvhat_l)
            dregions3 = eng.l2norm_bwd(regions3, vnorm, dvhat_l, None)
        if ctx.needs_input_grad[1]:
            dwords3 = eng.l2norm_bwd(words3, qnorm, dqhat, kq)
        return dregions3, dwords3, None, None, None, None, None, None


class LazyAttnMaps:
    """Stands in for the reference's ``attn_maps`` (losses.py:249): a list of B tensors (B, R, T) with the
    softmax-over-words probabilities of caption i against every image.  The reference materialises
    O(B^2 R T) floats whose only consumers are commented out (pretrain_DAMSM.py:221-228,
    trainer.py:446-453); here entry i is computed on first access."""

    def __init__(self, words3, regions3, mask_u8, engine):
        self._w, self._r, self._m, self._e = words3.detach(), regions3.detach(), mask_u8, engine
        self._cache = {}

    def __len__(self):
        return self._w.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        if i not in self._cache:
            b = self._r.shape[0]
            qhat, _, _, _ = self._e.l2norm_fwd(self._w[i:i + 1])
            vhat, _, _, _ = self._e.l2norm_fwd(self._r)
            q = qhat.expand(b, -1, -1).contiguous()
            m = self._m[i:i + 1].expand(b, -1).contiguous()
            _, attn, _ = self._e.func_attention_fwd(q, vhat, vhat, m, 1.0)
            self._cache[i] = attn.transpose(1, 2)            # (B, R, T) like losses.py:143-144
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _as_bnd(x: torch.Tensor, what: str):
    """(B, D, N) or (B, D, h, w) reference layout -> (B, N, D) strided view (no copy when possible)."""
    if x.dim() == 4:                                      # losses.py:350 hands over (B, 512, 7, 7)
        x = x.flatten(2)
    if x.dim() != 3:
        raise ValueError(f"{what} must be (B, D, N) or (B, D, h, w); got {tuple(x.shape)}")
    return x.permute(0, 2, 1)


def _mask_to_u8(words_mask, cap_lens, b, t, device):
    if words_mask is None:                                # stale 6-argument call sites (losses.py:352, trainer.py:235)
        if cap_lens is None:
            return torch.ones((b, t), dtype=torch.uint8, device=device)
        cl = torch.as_tensor(cap_lens).to(device=device, dtype=torch.int64).reshape(b, 1)
        return (torch.arange(t, device=device).reshape(1, t) < cl).to(torch.uint8)
    m = torch.as_tensor(words_mask)
    if m.dim() == 3:
        m = m.reshape(m.shape[0], -1)
    if tuple(m.shape) != (b, t):
        raise ValueError(f"words_mask must be ({b}, {t}); got {tuple(m.shape)}")
    return (m.to(device=device, non_blocking=True) != 0).to(torch.uint8).contiguous()


def _class_ids_tensor(class_ids, b, device):
    if class_ids is None:
        return None
    c = torch.as_tensor(np.asarray(class_ids) if not torch.is_tensor(class_ids) else class_ids)
    c = c.reshape(-1).to(device=device, dtype=torch.int64)
    if c.numel() != b:
        raise ValueError(f"class_ids must have {b} entries; got {c.numel()}")
    return c


TC_SMEM_LIMIT = 232448      # 227 KB of shared memory per CTA on sm_100


def tc_shape_supported(t, r, d):
    """True when the tcgen05 path covers (T, R, D): T <= 128, R <= 255, D a multiple of 64 and the resident caption
    tile + operand ring fit in shared memory (``damsm_words_tc_smem_bytes``)."""
    need = _lib.load().damsm_words_tc_smem_bytes(int(t), int(r), int(d))
    return 0 < need <= TC_SMEM_LIMIT


def _pick_precision(precision, t, r, d):
    if precision is None:
        return "fp32"
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    if precision == "bf16" and not tc_shape_supported(t, r, d):
        warnings.warn(f"words_loss: T={t}, R={r}, D={d} is outside the tensor-core kernel's range; "
                      "using the exact fp32 CUDA path", RuntimeWarning, stacklevel=3)
        return "fp32"
    return precision


def words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size,
               words_mask=None, gamma1=None, gamma2=None, gamma3=None, *, precision=None, group=None,
               engine=None):
    """Drop-in for ``miscc.losses.words_loss`` (losses.py:219).  Same positional signature, including the
    stale 6-argument form.  Extra keyword-only arguments: ``precision`` ('fp32' exact SIMT path, 'bf16'
    tcgen05 path), ``group`` (a torch.distributed process group: inputs are then this rank's shard of the
    batch and every rank's images serve as negatives), ``engine`` (tests only).

    Returns ``(loss0, loss1, attn_maps)``; ``match_labels=None`` gives ``(None, None, attn_maps)``.
    """
    regions3 = _as_bnd(region_features, "region_features")
    words3 = _as_bnd(words_embs, "words_embs")
    b, t, d = words3.shape
    if batch_size is not None and int(batch_size) != b:
        raise ValueError(f"batch_size={batch_size} does not match the tensors' batch {b}")
    if regions3.shape[2] != d:
        raise ValueError("words and regions must share the embedding size")
    eng = engine or get_engine(_pick_precision(precision, t, regions3.shape[1], d))
    dev = words3.device
    mask_u8 = _mask_to_u8(words_mask, cap_lens, b, t, dev)
    attn_maps = LazyAttnMaps(words3, regions3, mask_u8, eng)
    if match_labels is None:
        return None, None, attn_maps
    gammas = (DEFAULT_GAMMAS[0] if gamma1 is None else float(gamma1),
              DEFAULT_GAMMAS[1] if gamma2 is None else float(gamma2),
              DEFAULT_GAMMAS[2] if gamma3 is None else float(gamma3))
    world, _ = _world(group)
    labels = torch.as_tensor(match_labels).to(device=dev, dtype=torch.int64).contiguous()
    if labels.numel() != b * world:
        raise ValueError(f"match_labels must have {b * world} entries (global batch)")
    cls = _class_ids_tensor(class_ids, b, dev)
    loss0, loss1, _ = DamsmWordsLoss.apply(regions3, words3, mask_u8, labels, cls, gammas, eng, group)
    return loss0, loss1, attn_maps


# ----------------------------------------------------------------------------------------------- sentence loss
class DamsmSentLoss(torch.autograd.Function):
    """sent_loss (losses.py:51-91): rows = local image codes, columns = all caption codes."""

    @staticmethod
    def forward(ctx, img, txt, labels, cls_local, gamma3, eps, engine, group):
        world, rank = _world(group)
        bl = img.shape[0]
        b_total, row_offset = bl * world, rank * bl
        img = img.contiguous()
        txt_all = _all_gather_rows(txt.contiguous(), group)
        cls_all = _all_gather_rows(cls_local, group) if cls_local is not None else None
        # one launch each way where the loss is launch-bound (the reference's own batch sizes: measured 0.124 vs 0.145 ms
        # at B=48); from a few hundred rows on the tiled GEMM kernels of the unfused path win (B=512: 0.19 vs 0.61 ms)
        fused = (hasattr(engine, "sent_fwd") and engine.sent_fused_ok(gamma3)
                 and bl * txt_all.shape[0] <= SENT_FUSED_MAX_PAIRS)
        if fused:                                # norms + logits + class mask + row LSE + column partials
            logits, na, nb, row_lse, col_max, col_sum = engine.sent_fwd(img, txt_all, cls_local, cls_all, row_offset,
                                                                        gamma3, eps)
        else:
            logits, na, nb = engine.cos_logits(img, txt_all, gamma3, eps)
            row_lse, col_max, col_sum = engine.ce_stats(logits, cls_local, cls_all, row_offset)
        col_lse = combine_column_lse(col_max, col_sum, group)
        out2 = engine.ce_losses(logits, row_lse, col_lse, labels, row_offset, b_total)
        if group is not None:
            dist.all_reduce(out2, group=group)
        ctx.engine, ctx.group, ctx.gamma3, ctx.eps, ctx.fused = engine, group, gamma3, eps, fused
        ctx.row_offset, ctx.b_total = row_offset, b_total
        ctx.save_for_backward(img, txt_all, labels, na, nb, logits, row_lse, col_lse)
        return out2[0].clone(), out2[1].clone()

    @staticmethod
    def backward(ctx, g0, g1):
        img, txt_all, labels, na, nb, logits, row_lse, col_lse = ctx.saved_tensors
        gscale = torch.stack([g0.reshape(()), g1.reshape(())]).to(torch.float32)
        bwd = ctx.engine.sent_bwd if ctx.fused else ctx.engine.cos_logits_bwd
        da, db = bwd(img, txt_all, na, nb, logits, row_lse, col_lse, labels, gscale,
                     ctx.row_offset, ctx.b_total, ctx.gamma3, ctx.eps)
        dtxt = _reduce_scatter_rows(db, ctx.group) if ctx.needs_input_grad[1] else None
        return (da if ctx.needs_input_grad[0] else None), dtxt, None, None, None, None, None, None


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, *, gamma3=None, group=None,
              engine=None):
    """Drop-in for ``miscc.losses.sent_loss`` (losses.py:51).  ``gamma3`` defaults to the cfg value the
    reference reads at losses.py:79 (10.0 in every shipped yml)."""
    if cnn_code.dim() == 3:                                # (1, B, D) as produced by losses.py:69-71
        if cnn_code.shape[0] != 1:
            raise ValueError("sent_loss: only seq_len == 1 is meaningful (losses.py:82 squeezes it away)")
        cnn_code, rnn_code = cnn_code[0], rnn_code[0]
    if cnn_code.dim() != 2 or rnn_code.shape != cnn_code.shape:
        raise ValueError("sent_loss: cnn_code and rnn_code must both be (B, D)")
    b = cnn_code.shape[0]
    if batch_size is not None and int(batch_size) != b:
        raise ValueError(f"batch_size={batch_size} does not match the tensors' batch {b}")
    if labels is None:
        return None, None
    eng = engine or get_engine("fp32")
    dev = cnn_code.device
    world, _ = _world(group)
    lab = torch.as_tensor(labels).to(device=dev, dtype=torch.int64).contiguous()
    if lab.numel() != b * world:
        raise ValueError(f"labels must have {b * world} entries (global batch)")
    cls = _class_ids_tensor(class_ids, b, dev)
    g3 = DEFAULT_GAMMAS[2] if gamma3 is None else float(gamma3)
    out_dtype = cnn_code.dtype
    l0, l1 = DamsmSentLoss.apply(cnn_code.float(), rnn_code.float(), lab, cls, g3, float(eps), eng, group)
    return l0.to(out_dtype) if out_dtype != torch.float32 else l0, l1.to(out_dtype) if out_dtype != torch.float32 else l1


# ----------------------------------------------------------------------------------------------- NT-Xent
class DamsmNTXent(torch.autograd.Function):
    """NT_Xent.forward (nt_xent.py:16-35) with the mask of masks.py:3-17: one GEMM + one row kernel instead of
    the reference's (2B, 2B, D) broadcast."""

    @staticmethod
    def forward(ctx, z_i, z_j, inv_temp, eps, engine):
        z = torch.cat((z_i, z_j), dim=0).contiguous()                # nt_xent.py:22
        loss, sim, nrm, row_lse = engine.ntxent_fwd(z, inv_temp, eps)
        ctx.engine, ctx.inv_temp, ctx.eps, ctx.b = engine, inv_temp, eps, z_i.shape[0]
        ctx.save_for_backward(z, sim, nrm, row_lse)
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        z, sim, nrm, row_lse = ctx.saved_tensors
        gout = g.reshape(1).to(torch.float32).contiguous()
        dz = ctx.engine.ntxent_bwd(z, sim, nrm, row_lse, gout, ctx.inv_temp, ctx.eps)
        b = ctx.b
        return (dz[:b] if ctx.needs_input_grad[0] else None), (dz[b:] if ctx.needs_input_grad[1] else None), \
            None, None, None


def standard_ntxent_mask(batch_size):
    """The mask both reference builders produce (masks.py:3-17): False on the diagonal and on the +-B diagonals."""
    n2 = 2 * int(batch_size)
    idx = torch.arange(n2)
    same = idx.reshape(-1, 1) == idx.reshape(1, -1)
    pair = (idx.reshape(-1, 1) - idx.reshape(1, -1)).abs() == int(batch_size)
    return ~(same | pair)


def nt_xent(z_i, z_j, temperature, *, eps=1e-8, engine=None):
    """Functional form of ``NT_Xent(batch_size, temperature, mask, device)(z_i, z_j)`` (nt_xent.py:16-35)."""
    if z_i.dim() != 2 or z_j.shape != z_i.shape:
        raise ValueError("nt_xent: z_i and z_j must both be (B, D)")
    if not (float(temperature) > 0.0):
        raise ValueError("nt_xent: temperature must be positive")
    eng = engine or get_engine("fp32")
    out_dtype = z_i.dtype
    loss = DamsmNTXent.apply(z_i.float(), z_j.float(), 1.0 / float(temperature), float(eps), eng)
    return loss.to(out_dtype) if out_dtype != torch.float32 else loss


# ----------------------------------------------------------------------------------------------- R-precision
def r_precision_scores(img_code, sent_codes, eps=1e-8, *, engine=None):
    """The R-precision test of trainer.py:587-603 for a whole batch: ``img_code`` (B, D) generated-image codes,
    ``sent_codes`` (B, C, D) candidate sentence codes per image with the TRUE caption at index 0 (the reference
    concatenates it in front of 99 mismatched ones, :593).  Returns ``(scores0 (B, C), hit (B,) bool)`` with
    ``scores0 = img.sent / max(|img||sent|, eps)`` (:596-600) and ``hit = argmax(scores0) == 0`` (:601).  No gradient."""
    if img_code.dim() != 2 or sent_codes.dim() != 3 or sent_codes.shape[0] != img_code.shape[0] \
            or sent_codes.shape[2] != img_code.shape[1]:
        raise ValueError("r_precision_scores: img_code must be (B, D) and sent_codes (B, C, D)")
    eng = engine or get_engine("fp32")
    img = img_code.detach().float()
    cand = sent_codes.detach().float()
    img = img if img.stride(1) == 1 else img.contiguous()
    cand = cand if cand.stride(2) == 1 else cand.contiguous()
    scores, hit = eng.rprecision(img, cand, eps)
    return scores, hit.bool()


# ----------------------------------------------------------------------------------------------- generator_loss glue
class DamsmResizeNearest(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, hout, wout, engine):
        ctx.engine, ctx.in_hw, ctx.dtype = engine, tuple(x.shape[-2:]), x.dtype
        return engine.resize_nearest_fwd(x.contiguous(), hout, wout)

    @staticmethod
    def backward(ctx, dy):
        dx = ctx.engine.resize_nearest_bwd(dy.contiguous().float(), *ctx.in_hw)
        return dx.to(ctx.dtype), None, None, None


def clip_resize(images, size, *, engine=None):
    """Drop-in for ``F.interpolate(fake_imgs[i], size=image_size)`` at losses.py:348 (default mode 'nearest'): the
    256x256 generator output resized to CLIP's input resolution.  ``images`` (B, C, H, W), ``size`` int or (h, w).
    One gather kernel forward (bit-identical to torch), one scatter kernel backward."""
    if images.dim() < 3:
        raise ValueError("clip_resize: images must be (..., H, W)")
    hout, wout = (int(size), int(size)) if not isinstance(size, (tuple, list)) else (int(size[0]), int(size[1]))
    if images.element_size() not in (2, 4) or not images.is_floating_point():
        raise TypeError(f"clip_resize: unsupported dtype {images.dtype}")
    eng = engine or get_engine("fp32")
    return DamsmResizeNearest.apply(images, hout, wout, eng)


def generator_regions(region_features, grid=None):
    """``region_features[:, :, 1:].reshape(-1, D, g, g)`` of losses.py:350 WITHOUT the copy the reshape forces:
    ``region_features`` is the (B, D, R+1) permuted view ``encode_image_verbose`` returns (model.py:46-48; CLS token
    first); the result is a strided (B, D, g, g) view that ``words_loss`` reads in place."""
    if region_features.dim() != 3 or region_features.shape[2] < 2:
        raise ValueError("generator_regions: expected (B, D, R+1) with the CLS token first")
    r = region_features.shape[2] - 1
    g = int(math.isqrt(r)) if grid is None else int(grid)
    if g * g != r:
        raise ValueError(f"generator_regions: {r} regions are not a square grid")
    return region_features[:, :, 1:].unflatten(2, (g, g))


# ----------------------------------------------------------------------------------------------- region projection
class DamsmProjectRegions(torch.autograd.Function):
    """linear_subr + CLS drop (model.py:46,78; pretrain_DAMSM.py:125) with the l2norm prologue of words_loss fused into
    the GEMM epilogue.  Output: y (B, R, N) fp32."""

    @staticmethod
    def forward(ctx, subr, weight, bias, engine):
        if subr.dtype == torch.bfloat16:
            x, w = subr.contiguous(), weight.to(torch.bfloat16).contiguous()
        else:
            x, w = subr.float().contiguous(), weight.float().contiguous()
        b32 = None if bias is None else bias.float().contiguous()
        y, xhat, xhat16, norm, unorm = engine.project_regions_fwd(x, w, b32)
        ctx.engine = engine
        ctx.save_for_backward(x, w)
        ctx.dtypes = (subr.dtype, weight.dtype, None if bias is None else bias.dtype)
        _prologue_put(y, (xhat, xhat16, norm, unorm))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        need = ctx.needs_input_grad
        dx, dw, db = ctx.engine.project_regions_bwd(x.float(), w.float(), dy.float().contiguous(), need[0], need[1],
                                                    need[2] and ctx.dtypes[2] is not None)
        dt = ctx.dtypes
        return (None if dx is None else dx.to(dt[0]), None if dw is None else dw.to(dt[1]),
                None if db is None else db.to(dt[2]), None)


def project_regions(subr, weight, bias=None, *, engine=None):
    """``linear_subr(subr.view(-1, K)).view(B, -1, N)[:, 1:, :].permute(0, 2, 1)`` (model.py:46 / pretrain_DAMSM.py:359
    followed by :125): ``subr`` (B, R+1, K) ViT hidden states with the CLS token in row 0, ``weight`` (N, K), ``bias``
    (N).  Returns ``region_features`` (B, N, R) -- the tensor ``words_loss`` takes -- computed on the tensor cores
    (fp32 operands as TF32, bf16 operands as bf16; fp32 accumulate, fp32 result).  A ``words_loss`` call on the result
    reuses the normalised copies made in the GEMM epilogue."""
    if subr.dim() != 3 or weight.dim() != 2 or weight.shape[1] != subr.shape[2]:
        raise ValueError("project_regions: subr must be (B, R+1, K) and weight (N, K)")
    if subr.shape[1] < 2:
        raise ValueError("project_regions: need the CLS row plus at least one region")
    n = weight.shape[0]
    if n % 16 or not 16 <= n <= 512:
        raise ValueError("project_regions: N must be a multiple of 16 in [16, 512]")
    if bias is not None and bias.shape != (n,):
        raise ValueError("project_regions: bias must be (N,)")
    eng = engine or get_engine("bf16")
    return DamsmProjectRegions.apply(subr, weight, bias, eng).permute(0, 2, 1)


# ----------------------------------------------------------------------------------------------- rm_special_token
class DamsmRmSpecialToken(torch.autograd.Function):
    @staticmethod
    def forward(ctx, words_emb, mask_i64, engine):
        out, out_mask = engine.rm_special_token_fwd(words_emb, mask_i64)
        ctx.engine, ctx.n = engine, words_emb.shape[1]
        ctx.save_for_backward(mask_i64)
        ctx.mark_non_differentiable(out_mask)
        return out, out_mask

    @staticmethod
    def backward(ctx, dout, _dmask):
        (mask_i64,) = ctx.saved_tensors
        return ctx.engine.rm_special_token_bwd(dout.contiguous(), mask_i64, ctx.n), None, None


def rm_special_token(mask, words_emb, *, engine=None):
    """Drop-in for ``rm_special_token`` (pretrain_DAMSM.py:58-79): removes the <sos> row and the <eos> row (the one
    before the first 0 of the attention mask; the last row when the mask has no 0) from ``words_emb`` (B, n, D) and
    ``mask`` (B, n); returns ``(words_emb_new (B, n-2, D), mask_new (B, n-2))``.  One gather kernel, no host sync.
    A caption whose mask starts with fewer than two 1s makes the reference's ``torch.stack`` fail; here it is treated
    as <sos><eos> only."""
    if words_emb.dim() != 3 or mask.dim() != 2 or mask.shape != words_emb.shape[:2]:
        raise ValueError("rm_special_token: words_emb must be (B, n, D) and mask (B, n)")
    if words_emb.shape[1] < 3:
        raise ValueError("rm_special_token: need at least 3 tokens per caption")
    if words_emb.element_size() not in (2, 4):
        raise TypeError(f"rm_special_token: unsupported dtype {words_emb.dtype}")
    eng = engine or get_engine("fp32")
    x = words_emb if words_emb.stride(2) == 1 else words_emb.contiguous()
    m64 = mask.to(device=words_emb.device, dtype=torch.int64)
    out, out_mask = DamsmRmSpecialToken.apply(x, m64, eng)
    return out, (out_mask if mask.dtype == torch.int64 else out_mask.to(mask.dtype))


# ----------------------------------------------------------------------------------------------- func_attention
class DamsmFuncAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query3, context3, mask_u8, gamma1, engine):
        qhat, _, qnorm, _ = engine.l2norm_fwd(query3)
        vhat, _, vnorm, _ = engine.l2norm_fwd(context3)
        ctx32 = context3 if context3.dtype == torch.float32 else context3.float()
        wc, attn, attn2 = engine.func_attention_fwd(qhat, vhat, ctx32, mask_u8, gamma1)
        ctx.engine, ctx.gamma1 = engine, gamma1
        ctx.save_for_backward(query3, context3, ctx32, qhat, vhat, qnorm, vnorm, attn, attn2)
        return wc, attn

    @staticmethod
    def backward(ctx, d_wc, d_attn):
        query3, context3, ctx32, qhat, vhat, qnorm, vnorm, attn, attn2 = ctx.saved_tensors
        eng = ctx.engine
        d_wc = d_wc.contiguous().float() if d_wc is not None else None
        d_attn = d_attn.contiguous().float() if d_attn is not None else None
        dqhat, dvhat, dctx = eng.func_attention_bwd(qhat, vhat, ctx32, attn, attn2, d_wc, d_attn, ctx.gamma1)
        dq = eng.l2norm_bwd(query3, qnorm, dqhat, None) if ctx.needs_input_grad[0] else None
        dc = None
        if ctx.needs_input_grad[1]:
            dc = eng.l2norm_bwd(context3, vnorm, dvhat, None) + dctx.to(context3.dtype)
        return dq, dc, None, None, None


def func_attention(query, context, gamma1, query_mask, *, engine=None):
    """Drop-in for ``GlobalAttention.func_attention`` (GlobalAttention.py:38).

    query (B, D, T), context (B, D, R) with R a perfect square (:54), query_mask (B, 1, T).
    Returns (weightedContext (B, T, D), attn (B, T, sqrt R, sqrt R))."""
    if query.dim() != 3 or context.dim() != 3:
        raise ValueError("func_attention: query must be (B, D, T) and context (B, D, R)")
    b, d, t = query.shape
    r = context.shape[2]
    side = int(math.sqrt(r))
    if side * side != r:
        raise ValueError(f"func_attention: number of regions {r} is not a perfect square (GlobalAttention.py:54,156)")
    eng = engine or get_engine("fp32")
    m = torch.as_tensor(query_mask)
    mask_u8 = (m.reshape(b, t).to(query.device) != 0).to(torch.uint8).contiguous()
    wc, attn = DamsmFuncAttention.apply(query.permute(0, 2, 1), context.permute(0, 2, 1), mask_u8, float(gamma1), eng)
    if query.dtype != torch.float32:
        wc, attn = wc.to(query.dtype), attn.to(query.dtype)
    return wc, attn.view(b, t, side, side)
