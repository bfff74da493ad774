"""Block-level compute engine: thin, typed wrappers over the C ABI (include/damsm_b200.h).

``CudaEngine`` is the product engine -- every method launches hand-written sm_100a kernels from
libdamsm_b200.so on the current CUDA stream and raises if that is not possible.  The host logic in
``ops.py`` (autograd plumbing, sharding, collectives) only talks to this small interface, which is what
lets the CPU-only ``gloo`` tests exercise the multi-rank logic with a checker engine injected from
``tests/`` -- the package itself contains no CPU implementation.
"""
from __future__ import annotations

import torch

from . import _lib

_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.DamsmError(
                "DAMSM kernels are CUDA-only (sm_100a); got a CPU tensor and there is no CPU fallback")


def _on_tensor_device(fn):
    """Run an engine method with the device of its first CUDA tensor argument current: the C ABI launches on the
    CURRENT device's stream, so a caller that holds tensors of another device (``cuda:1`` while ``cuda:0`` is
    current) would otherwise launch on the wrong GPU."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kw):
        for a in list(args) + list(kw.values()):
            if isinstance(a, dict):
                a = next((v for v in a.values() if torch.is_tensor(v) and v.is_cuda), None)
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(self, *args, **kw)
        return fn(self, *args, **kw)
    return wrapped


class _EngineMeta(type):
    def __new__(mcls, name, bases, ns):
        for k, v in list(ns.items()):
            if callable(v) and not k.startswith("_") or k == "_words_bwd_tc":
                ns[k] = _on_tensor_device(v)
        return super().__new__(mcls, name, bases, ns)


class CudaEngine(metaclass=_EngineMeta):
    """Exact-fp32 SIMT path (precision='fp32') and bf16 tcgen05 path (precision='bf16')."""

    name = "cuda"

    def __init__(self, precision: str = "fp32"):
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"unknown precision {precision!r}")
        self.precision = precision
        _lib.load()

    # ---- l2norm ---------------------------------------------------------------------------------------
    def l2norm_fwd(self, x3: torch.Tensor, want_bf16: bool = False, pad8: bool = False):
        """x3: (nb, nv, D) strided view.  Returns (xhat_f32, xhat_16|None, norm, unorm).  ``want_bf16`` asks for
        the 16-bit operand copy of the tensor-core path (stored as fp16: |xhat| <= 1, so fp16 keeps 3 more
        mantissa bits than bf16 at the same tensor-core rate); with ``pad8`` its vector count is padded to a
        multiple of 8 with zero rows."""
        _require_cuda(x3)
        if x3.dtype not in _DTYPE_CODE:
            raise TypeError(f"unsupported dtype {x3.dtype}")
        nb, nv, d = x3.shape
        dev = x3.device
        xhat = torch.empty((nb, nv, d), device=dev, dtype=torch.float32)
        nv_pad = (nv + 7) // 8 * 8 if pad8 else nv
        xhat16 = torch.empty((nb, nv_pad, d), device=dev, dtype=torch.float16) if want_bf16 else None
        norm = torch.empty((nb, nv), device=dev, dtype=torch.float32)
        unorm = torch.empty((nb, nv), device=dev, dtype=torch.float32)
        sb, sv, sd = x3.stride()
        _lib.call("damsm_l2norm_fwd", x3.data_ptr(), _DTYPE_CODE[x3.dtype], nb, nv, d, sb, sv, sd,
                  xhat.data_ptr(), _lib.ptr(xhat16), nv_pad, norm.data_ptr(), unorm.data_ptr(), _stream())
        return xhat, xhat16, norm, unorm

    def l2norm_bwd(self, x3: torch.Tensor, norm, dxhat, kq=None):
        """Gradient w.r.t. the raw x3, returned as a tensor with x3's shape and dtype whose memory
        layout matches x3's (so autograd can hand it straight back through the permute)."""
        _require_cuda(x3, norm, dxhat, kq)
        nb, nv, d = x3.shape
        dx = torch.empty_strided(x3.shape, _dense_strides_like(x3), device=x3.device, dtype=x3.dtype)
        sb, sv, sd = x3.stride()
        dsb, dsv, dsd = dx.stride()
        if dxhat.stride(2) != 1:
            dxhat = dxhat.contiguous()
        _lib.call("damsm_l2norm_bwd", x3.data_ptr(), _DTYPE_CODE[x3.dtype], nb, nv, d, sb, sv, sd,
                  norm.data_ptr(), dxhat.data_ptr(), dxhat.stride(0), dxhat.stride(1), _lib.ptr(kq),
                  dx.data_ptr(), dsb, dsv, dsd, _stream())
        return dx

    # ---- word / region scores ---------------------------------------------------------------------------
    def gram(self, vhat: torch.Tensor):
        """Per-image Gram matrices G_j = vhat_j vhat_j^T (bc,R,R) fp32 of the LOCAL images."""
        bc, r, d = vhat.shape
        gram = torch.empty((bc, r, r), device=vhat.device, dtype=torch.float32)
        _lib.call("damsm_gram_f32", vhat.data_ptr(), bc, r, d, gram.data_ptr(), _stream())
        return gram

    def pack_columns(self, gram: torch.Tensor, vhat, vhat16=None):
        """Image-side operands of the pair kernels from the (gathered) Gram matrices and normalised regions:
        fp32 path: gram + vhat; tensor-core path: the padded fp16 Gram form with the appended row of ones + vhat16."""
        col = {"gram": gram, "shape": tuple(gram.shape[:2])}
        if self.precision == "bf16":
            col["gx"] = self._pack_gx(gram)
            col["vhat16"] = vhat16
        return col

    def _pack_gx(self, gram):
        bc, r, _ = gram.shape
        rk = _lib.load().damsm_words_tc_gx_cols(r)
        gx = torch.empty((bc, r + 1, rk), device=gram.device, dtype=torch.float16)
        _lib.call("damsm_gram_pack_tc", gram.data_ptr(), bc, r, gx.data_ptr(), _stream())
        return gx

    def image_side(self, vhat_l, vhat16_l, gather):
        """Image-side operands of the pair kernels for ALL ranks' images from this rank's normalised regions: the Gram
        matrices are formed (and, tensor-core path, packed to the padded fp16 form with the row of ones) for the LOCAL
        images only; ``gather`` (all-gather over the ranks, identity on one GPU) then moves exactly what the pair
        kernels read -- fp16 gx + fp16 vhat (tensor-core path) or fp32 gram + fp32 vhat (exact path)."""
        gram_l = self.gram(vhat_l)
        r = gram_l.shape[1]
        if self.precision == "bf16":
            gx = gather(self._pack_gx(gram_l))
            return {"gram": None, "shape": (gx.shape[0], r), "gx": gx, "vhat16": gather(vhat16_l)}, vhat_l
        gram = gather(gram_l)
        return {"gram": gram, "shape": (gram.shape[0], r)}, gather(vhat_l)

    def words_prepare_columns(self, vhat: torch.Tensor, vhat16=None):
        """Single-GPU convenience: gram + pack_columns."""
        return self.pack_columns(self.gram(vhat), vhat, vhat16)

    def gram_bwd(self, hmat, vhat, dvhat):
        """dvhat_j -= H_j vhat_j (in place) for the LOCAL images, after H and dvhat have been reduced over ranks."""
        bc, r, d = vhat.shape
        _lib.call("damsm_gram_bwd_f32", hmat.data_ptr(), vhat.data_ptr(), bc, r, d, dvhat.data_ptr(), _stream())
        return dvhat

    def words_fwd(self, qhat, qhat16, vhat, col, unorm, mask_u8, gammas, want_stats=True):
        _require_cuda(qhat, unorm, mask_u8)
        br, t, d = qhat.shape
        bc, r = col["shape"]
        sim = torch.empty((br, bc), device=qhat.device, dtype=torch.float32)
        if self.precision == "bf16":
            # per-pair, per-word scalars (rho, ||c||, 1/Y) for the backward: 12*T bytes per pair
            stats = torch.empty((br, bc, 3, t), device=qhat.device, dtype=torch.float32) if want_stats else None
            col["stats"] = stats
            plan = self.words_plan(mask_u8, t)
            pad = self.pad_terms_fwd(col["vhat16"], qhat16, unorm, plan["nw"], t, float(gammas[1])) if t > 16 else None
            col["plan"], col["pad"] = plan, pad
            _lib.call("damsm_words_fwd_tc", qhat16.data_ptr(), qhat16.shape[1], col["vhat16"].data_ptr(),
                      col["gx"].data_ptr(),
                      unorm.data_ptr(), mask_u8.data_ptr(), plan["nw"].data_ptr(), plan["order"].data_ptr(),
                      _lib.ptr(pad["epad"]) if pad else None, br, bc, t, r, d,
                      float(gammas[0]), float(gammas[1]), float(gammas[2]), sim.data_ptr(), _lib.ptr(stats),
                      _stream())
            _lib.add_launches((t > 32) + (t > 64))        # one launch per caption-length group (nw <= 32, <= 64, longer)
        else:
            _lib.call("damsm_words_fwd_f32", qhat.data_ptr(), vhat.data_ptr(), col["gram"].data_ptr(),
                      unorm.data_ptr(), mask_u8.data_ptr(), br, bc, t, r, d,
                      float(gammas[0]), float(gammas[1]), float(gammas[2]), sim.data_ptr(), _stream())
        return sim

    # ---- skip-padded-words plan and closed form (tensor-core path) -----------------------------------------------
    def words_plan(self, mask_u8, t):
        """Per-caption word counts ``nw`` (multiples of 16 covering the last unmasked word) and the captions sorted by
        them, on the device; a pinned host copy of ``nw`` is started for the backward (which sizes its launches by it)."""
        br = mask_u8.shape[0]
        dev = mask_u8.device
        nw = torch.empty(br, device=dev, dtype=torch.int32)
        order = torch.empty(br, device=dev, dtype=torch.int32)
        _lib.call("damsm_words_tc_plan", mask_u8.data_ptr(), br, t, nw.data_ptr(), order.data_ptr(), _stream())
        nw_host = torch.empty(br, dtype=torch.int32, pin_memory=True)
        nw_host.copy_(nw, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return dict(nw=nw, order=order, nw_host=nw_host, event=ev)

    def pad_terms_fwd(self, vhat16, qhat16, unorm, nw, t, gamma2):
        """Mean region per image and epad[i][j] = sum over the skipped words of exp(gamma2 rho_bar) (pad_terms.cu)."""
        bc, r, d = vhat16.shape
        br, tp, _ = qhat16.shape
        dev = vhat16.device
        out = dict(vbar32=torch.empty((bc, d), device=dev, dtype=torch.float32),
                   vbar16=torch.empty((bc, d), device=dev, dtype=torch.float16),
                   nbar=torch.empty(bc, device=dev, dtype=torch.float32),
                   rn=torch.empty(bc, device=dev, dtype=torch.float32),
                   epad=torch.empty((br, bc), device=dev, dtype=torch.float32))
        _lib.call("damsm_pad_terms_fwd", vhat16.data_ptr(), qhat16.data_ptr(), unorm.data_ptr(), nw.data_ptr(),
                  br, bc, t, tp, r, d, float(gamma2), out["vbar32"].data_ptr(), out["vbar16"].data_ptr(),
                  out["nbar"].data_ptr(), out["rn"].data_ptr(), out["epad"].data_ptr(), _stream())
        return out

    def words_bwd(self, qhat, qhat16, vhat, col, unorm, mask_u8, sim, row_lse, col_lse, labels, gscale,
                  row_offset, b_total, gammas, need_dq=True, need_dv=True):
        """Returns (dqhat (br,T,D), dvhat (bc,R,D), hmat (bc,R,R), kq (br,T)); dvhat and hmat are partial sums over
        this rank's caption rows and dvhat does not yet contain the -H vhat term (see gram_bwd).  ``vhat`` is only
        read by the fp32 path.  Tensor-core path: ``need_dq`` / ``need_dv`` = False skip that side's GEMMs and return
        None for it (the fp32 path always computes both)."""
        gram = col["gram"]
        br, t, d = qhat.shape
        bc, r = col["shape"]
        dev = qhat.device
        if self.precision == "bf16":
            return self._words_bwd_tc(qhat, qhat16, col, unorm, mask_u8, sim, row_lse, col_lse, labels, gscale,
                                      row_offset, b_total, gammas, br, bc, t, r, d, need_dq, need_dv)
        dqhat = torch.zeros((br, t, d), device=dev, dtype=torch.float32)
        dvhat = torch.zeros((bc, r, d), device=dev, dtype=torch.float32)
        hmat = torch.zeros((bc, r, r), device=dev, dtype=torch.float32)
        kq = torch.zeros((br, t), device=dev, dtype=torch.float32)
        _lib.call("damsm_words_bwd_f32", qhat.data_ptr(), vhat.data_ptr(), gram.data_ptr(), unorm.data_ptr(),
                  mask_u8.data_ptr(), sim.data_ptr(), row_lse.data_ptr(), col_lse.data_ptr(), _lib.ptr(labels),
                  gscale.data_ptr(), int(row_offset), int(b_total), br, bc, t, r, d,
                  float(gammas[0]), float(gammas[1]), float(gammas[2]),
                  dqhat.data_ptr(), dvhat.data_ptr(), hmat.data_ptr(), kq.data_ptr(), _stream())
        return dqhat, dvhat, hmat, kq

    # scratch for the tensor-core backward: the fused kernel + GEMMs run chunk by chunk inside it.  Larger chunks
    # mean fewer launches and a longer K for the gradient GEMMs; default = a third of the HBM that is free at the
    # FIRST call with a given shape (at least 6 GiB, never more than 80 % of what is free), remembered per shape so
    # that the chunking -- and with it the launch geometry -- does not drift from iteration to iteration.
    tc_workspace_bytes = None
    _tc_ws_cols = {}

    def _tc_workspace_cols(self, dev, br, bc, t, r):
        """Scratch columns (= words of captions) one chunk may hold."""
        lib = _lib.load()
        col_bytes = lib.damsm_words_bwd_tc_col_bytes(bc, r)
        fixed = lib.damsm_words_bwd_tc_fixed_bytes()
        key = (dev.index, br, bc, t, r, self.tc_workspace_bytes)
        cols = self._tc_ws_cols.get(key)
        if cols is None:
            budget = self.tc_workspace_bytes
            if budget is None:
                free = torch.cuda.mem_get_info(dev)[0]
                budget = min(max(6 << 30, free // 3), int(free * 0.8))
            cols = int(max(128, (budget - fixed) // col_bytes // 16 * 16))
            self._tc_ws_cols[key] = cols
        return cols, col_bytes, fixed

    @staticmethod
    def _tc_chunks(nw_sorted, kc_max):
        """Chunk boundaries (positions in the sorted caption order) with at most kc_max scratch columns per chunk and
        about equal column counts; koff = prefix sums of the sorted word counts."""
        import numpy as np
        koff = np.zeros(len(nw_sorted) + 1, dtype=np.int64)
        np.cumsum(nw_sorted, out=koff[1:])
        total = int(koff[-1])
        n = len(nw_sorted)
        if n and int(nw_sorted.max()) > kc_max:
            raise _lib.DamsmError("tensor-core backward: workspace smaller than one caption")
        n_chunks = max(1, -(-total // kc_max))
        while True:
            # boundaries at equal quantiles of the column count; more chunks if one of them does not fit
            cuts = [int(np.searchsorted(koff, total * c / n_chunks, side="left")) for c in range(1, n_chunks)]
            pos = sorted(set([0] + [min(max(c, 1), n) for c in cuts] + [n]))
            sizes = koff[pos[1:]] - koff[pos[:-1]]
            if sizes.max() <= kc_max or n_chunks >= n:
                break
            n_chunks += 1
        return koff, np.asarray(pos, dtype=np.int64)

    def _words_bwd_tc(self, qhat, qhat16, col, unorm, mask_u8, sim, row_lse, col_lse, labels, gscale,
                      row_offset, b_total, gammas, br, bc, t, r, d, need_dq=True, need_dv=True):
        import numpy as np
        dev = qhat16.device
        tp = qhat16.shape[1]
        plan, pad = col["plan"], col.get("pad")
        plan["event"].synchronize()                         # the pinned copy of nw (started before the forward kernel)
        nw_host = plan["nw_host"].numpy()
        order_host = np.argsort(-nw_host.astype(np.int64), kind="stable").astype(np.int32)
        kc_max, col_bytes, fixed = self._tc_workspace_cols(dev, br, bc, t, r)
        koff_host, chunk_pos = self._tc_chunks(nw_host[order_host].astype(np.int64), kc_max)
        total_k = int(koff_host[-1])
        kc_need = int((koff_host[chunk_pos[1:]] - koff_host[chunk_pos[:-1]]).max())
        ws_bytes = fixed + kc_need * col_bytes
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        order = torch.from_numpy(order_host).to(dev)
        koff = torch.from_numpy(koff_host).to(dev)
        qpack16 = torch.empty((total_k, d), device=dev, dtype=torch.float16)
        dqpack = torch.empty((total_k, d), device=dev, dtype=torch.float32) if need_dq else None
        dqhat = torch.empty((br, tp, d), device=dev, dtype=torch.float32) if need_dq else None
        dvhat = torch.zeros((bc, r, d), device=dev, dtype=torch.float32) if need_dv else None
        hmat = torch.zeros((bc, r, r), device=dev, dtype=torch.float32) if need_dv else None
        kq = torch.zeros((br, t), device=dev, dtype=torch.float32)
        n_chunks = len(chunk_pos) - 1
        _lib.call("damsm_words_bwd_tc", qhat16.data_ptr(), tp, col["vhat16"].data_ptr(), col["gx"].data_ptr(),
                  unorm.data_ptr(), mask_u8.data_ptr(), plan["nw"].data_ptr(), order.data_ptr(), koff.data_ptr(),
                  koff_host.ctypes.data, chunk_pos.ctypes.data, n_chunks, sim.data_ptr(), col["stats"].data_ptr(),
                  row_lse.data_ptr(), col_lse.data_ptr(),
                  _lib.ptr(labels), gscale.data_ptr(), int(row_offset), int(b_total), br, bc, t, r, d,
                  float(gammas[0]), float(gammas[1]), float(gammas[2]), ws.data_ptr(), ws_bytes,
                  qpack16.data_ptr(), _lib.ptr(dqpack), _lib.ptr(dvhat), _lib.ptr(hmat), kq.data_ptr(), _stream())
        # own kernels: scalars + packing (counted by the call) + per chunk the fused recompute, the dvhat GEMM and the
        # H kernel (image side), the dqhat GEMM (caption side)
        # the fused kernel is launched once per caption-length group (nw <= 32, <= 64, longer) present in a chunk
        nws = nw_host[order_host].astype(np.int64)
        gid = (nws > 32).astype(np.int64) + (nws > 64)
        n_fused = sum(len(np.unique(gid[chunk_pos[c]:chunk_pos[c + 1]])) for c in range(n_chunks))
        _lib.add_launches(n_fused + ((2 if need_dv else 0) + (1 if need_dq else 0)) * n_chunks)
        # closed form of the skipped words + unpacking of the packed word-row gradients (pad_terms.cu)
        coef = dqpad = dvbar = scal = None
        if pad is not None:
            coef = torch.empty((bc, br * tp), device=dev, dtype=torch.float16)
            dqpad = torch.empty((br * tp, d), device=dev, dtype=torch.float32) if need_dq else None
            dvbar = torch.empty((bc, d), device=dev, dtype=torch.float32) if need_dv else None
            scal = torch.empty(64, device=dev, dtype=torch.float32)
        g = lambda k: _lib.ptr(pad[k]) if pad is not None else None
        _lib.call("damsm_pad_terms_bwd", g("vbar32"), g("vbar16"), g("nbar"), g("rn"), qhat16.data_ptr(), qhat.data_ptr(),
                  unorm.data_ptr(), plan["nw"].data_ptr(), order.data_ptr(), koff.data_ptr(), sim.data_ptr(),
                  row_lse.data_ptr(), col_lse.data_ptr(), _lib.ptr(labels), gscale.data_ptr(), int(row_offset),
                  int(b_total), br, bc, t, tp, r, d, float(gammas[1]), float(gammas[2]), _lib.ptr(coef), _lib.ptr(dqpad),
                  _lib.ptr(dvbar), _lib.ptr(scal), _lib.ptr(dqpack), _lib.ptr(dqhat), kq.data_ptr(), _lib.ptr(dvhat),
                  _stream())
        if pad is not None:
            _lib.add_launches(2 + (2 if need_dq else 0) + (2 if need_dv else 0))
        elif need_dq:
            _lib.add_launches(1)
        return (dqhat[:, :t, :] if need_dq else None), dvhat, hmat, kq

    # ---- dense contraction on the tensor cores (gemm_tc.cu) ----------------------------------------------------
    def gemm_tc(self, a, b, a_mn=False, b_mn=False, out=None, accumulate=False, alpha=1.0, alpha_dev=None):
        """C (M, N) fp32 =|+= alpha * A . B.  ``a``: (M, K) row-major, or with ``a_mn`` (K, M) row-major (A^T as it
        lies); ``b``: (N, K) row-major, or with ``b_mn`` (K, N) row-major.  fp16 / bf16 operands, or fp32 (as TF32)."""
        _require_cuda(a, b, out, alpha_dev)
        if a.dtype != b.dtype or a.dtype not in _DTYPE_CODE:
            raise TypeError("gemm_tc: operands must share a dtype (fp16, bf16 or fp32)")
        if a.dim() != 2 or b.dim() != 2 or a.stride(1) != 1 or b.stride(1) != 1:
            raise ValueError("gemm_tc: operands must be 2-D with a contiguous last dimension")
        m, k = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
        n, kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
        if kb != k:
            raise ValueError(f"gemm_tc: K mismatch ({k} vs {kb})")
        if out is None:
            if accumulate:
                raise ValueError("gemm_tc: accumulate needs an output tensor")
            out = torch.empty((m, n), device=a.device, dtype=torch.float32)
        fmt = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[a.dtype]
        _lib.call("damsm_gemm_tc", a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn), fmt,
                  m, n, k, float(alpha), _lib.ptr(alpha_dev), int(accumulate), out.data_ptr(), out.stride(0), _stream())
        return out

    # ---- masked bidirectional cross-entropy ---------------------------------------------------------------
    def ce_stats(self, logits, cls_rows, cls_cols, row_offset):
        """Masks ``logits`` in place; returns (row_lse, col_max, col_sum)."""
        _require_cuda(logits, cls_rows, cls_cols)
        br, bc = logits.shape
        dev = logits.device
        row_lse = torch.empty(br, device=dev, dtype=torch.float32)
        col_max = torch.empty(bc, device=dev, dtype=torch.float32)
        col_sum = torch.empty(bc, device=dev, dtype=torch.float32)
        _lib.call("damsm_ce_stats_f32", logits.data_ptr(), _lib.ptr(cls_rows), _lib.ptr(cls_cols), int(row_offset),
                  br, bc, row_lse.data_ptr(), col_max.data_ptr(), col_sum.data_ptr(), _stream())
        return row_lse, col_max, col_sum

    def ce_losses(self, logits, row_lse, col_lse, labels, row_offset, b_total):
        br, bc = logits.shape
        out = torch.empty(2, device=logits.device, dtype=torch.float32)
        _lib.call("damsm_ce_losses_f32", logits.data_ptr(), row_lse.data_ptr(), col_lse.data_ptr(),
                  _lib.ptr(labels), int(row_offset), br, bc, int(b_total), out.data_ptr(), _stream())
        return out

    # ---- sentence logits ------------------------------------------------------------------------------------
    def cos_logits(self, a, b, gamma3, eps):
        _require_cuda(a, b)
        br, d = a.shape
        bc = b.shape[0]
        dev = a.device
        logits = torch.empty((br, bc), device=dev, dtype=torch.float32)
        na = torch.empty(br, device=dev, dtype=torch.float32)
        nb = torch.empty(bc, device=dev, dtype=torch.float32)
        _lib.call("damsm_cos_logits_f32", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), br, bc, d,
                  float(gamma3), float(eps), logits.data_ptr(), na.data_ptr(), nb.data_ptr(), _stream())
        return logits, na, nb

    def cos_logits_bwd(self, a, b, na, nb, logits, row_lse, col_lse, labels, gscale, row_offset, b_total,
                       gamma3, eps):
        br, d = a.shape
        bc = b.shape[0]
        dev = a.device
        work = torch.empty(br * bc + br + bc, device=dev, dtype=torch.float32)
        da = torch.empty((br, d), device=dev, dtype=torch.float32)
        db = torch.empty((bc, d), device=dev, dtype=torch.float32)
        _lib.call("damsm_cos_logits_bwd_f32", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                  na.data_ptr(), nb.data_ptr(), logits.data_ptr(), row_lse.data_ptr(), col_lse.data_ptr(),
                  _lib.ptr(labels), gscale.data_ptr(), int(row_offset), int(b_total), br, bc, d,
                  float(gamma3), float(eps), work.data_ptr(), da.data_ptr(), db.data_ptr(), _stream())
        return da, db

    # ---- sentence loss, one launch each way (sent_fused.cu) -------------------------------------------------------
    def sent_fused_ok(self, gamma3):
        return bool(_lib.load().damsm_sent_fused_ok(float(gamma3)))

    def sent_fwd(self, a, b, cls_rows, cls_cols, row_offset, gamma3, eps):
        """Cosine logits + class mask + row LSE + column partials in ONE kernel.  Returns (logits, na, nb, row_lse,
        col_max, col_sum) with (col_max, col_sum) in the partial form of ce_stats (col_max = 0)."""
        _require_cuda(a, b, cls_rows, cls_cols)
        br, d = a.shape
        bc = b.shape[0]
        dev = a.device
        logits = torch.empty((br, bc), device=dev, dtype=torch.float32)
        na = torch.empty(br, device=dev, dtype=torch.float32)
        nb = torch.empty(bc, device=dev, dtype=torch.float32)
        row_lse = torch.empty(br, device=dev, dtype=torch.float32)
        col_max = torch.empty(bc, device=dev, dtype=torch.float32)
        col_sum = torch.empty(bc, device=dev, dtype=torch.float32)
        _lib.call("damsm_sent_fwd_fused_f32", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), _lib.ptr(cls_rows),
                  _lib.ptr(cls_cols), int(row_offset), br, bc, d, float(gamma3), float(eps), logits.data_ptr(),
                  na.data_ptr(), nb.data_ptr(), row_lse.data_ptr(), col_max.data_ptr(), col_sum.data_ptr(), _stream())
        return logits, na, nb, row_lse, col_max, col_sum

    def sent_bwd(self, a, b, na, nb, logits, row_lse, col_lse, labels, gscale, row_offset, b_total, gamma3, eps):
        br, d = a.shape
        bc = b.shape[0]
        da = torch.empty((br, d), device=a.device, dtype=torch.float32)
        db = torch.empty((bc, d), device=a.device, dtype=torch.float32)
        _lib.call("damsm_sent_bwd_fused_f32", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), na.data_ptr(),
                  nb.data_ptr(), logits.data_ptr(), row_lse.data_ptr(), col_lse.data_ptr(), _lib.ptr(labels),
                  gscale.data_ptr(), int(row_offset), int(b_total), br, bc, d, float(gamma3), float(eps),
                  da.data_ptr(), db.data_ptr(), _stream())
        return da, db

    # ---- NT-Xent (nt_xent.py) -------------------------------------------------------------------------------
    def ntxent_fwd(self, z, inv_temp, eps):
        _require_cuda(z)
        n2, d = z.shape
        dev = z.device
        sim = torch.empty((n2, n2), device=dev, dtype=torch.float32)
        nrm = torch.empty(n2, device=dev, dtype=torch.float32)
        row_lse = torch.empty(n2, device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        _lib.call("damsm_ntxent_fwd_f32", z.data_ptr(), z.stride(0), n2, d, float(inv_temp), float(eps),
                  sim.data_ptr(), nrm.data_ptr(), row_lse.data_ptr(), loss.data_ptr(), _stream())
        return loss, sim, nrm, row_lse

    def ntxent_bwd(self, z, sim, nrm, row_lse, gout, inv_temp, eps):
        n2, d = z.shape
        work = torch.empty(n2 * n2 + n2, device=z.device, dtype=torch.float32)
        dz = torch.empty((n2, d), device=z.device, dtype=torch.float32)
        _lib.call("damsm_ntxent_bwd_f32", z.data_ptr(), z.stride(0), n2, d, float(inv_temp), float(eps),
                  sim.data_ptr(), nrm.data_ptr(), row_lse.data_ptr(), gout.data_ptr(), work.data_ptr(),
                  dz.data_ptr(), _stream())
        return dz

    # ---- R-precision scoring (trainer.py:587-603) ---------------------------------------------------------------
    def rprecision(self, img, cand, eps):
        _require_cuda(img, cand)
        b, c, d = cand.shape
        scores = torch.empty((b, c), device=img.device, dtype=torch.float32)
        hit = torch.empty(b, device=img.device, dtype=torch.int32)
        _lib.call("damsm_rprecision_f32", img.data_ptr(), img.stride(0), cand.data_ptr(), cand.stride(0), cand.stride(1),
                  b, c, d, float(eps), scores.data_ptr(), hit.data_ptr(), _stream())
        return scores, hit

    # ---- region projection fused with the l2norm prologue (model.py:46,78) ----------------------------------------
    def project_regions_fwd(self, x, w, bias):
        """x (B, R+1, K) contiguous fp32 or bf16 (row 0 of every image = CLS), w (N, K) same dtype, bias (N) fp32 or
        None.  Returns y (B,R,N) fp32, xhat, xhat16 (fp16), norm (B,R), unorm (B,R) -- what l2norm_fwd would give."""
        _require_cuda(x, w, bias)
        b, rp1, k = x.shape
        n = w.shape[0]
        r = rp1 - 1
        dev = x.device
        y = torch.empty((b, r, n), device=dev, dtype=torch.float32)
        xhat = torch.empty((b, r, n), device=dev, dtype=torch.float32)
        xhat16 = torch.empty((b, r, n), device=dev, dtype=torch.float16)
        norm = torch.empty((b, r), device=dev, dtype=torch.float32)
        unorm = torch.empty((b, r), device=dev, dtype=torch.float32)
        _lib.call("damsm_project_regions_fwd", x.data_ptr(), 0 if x.dtype == torch.float32 else 1, b, r, k, w.data_ptr(),
                  _lib.ptr(bias), n, y.data_ptr(), xhat.data_ptr(), xhat16.data_ptr(), norm.data_ptr(), unorm.data_ptr(),
                  _stream())
        return y, xhat, xhat16, norm, unorm

    def project_regions_bwd(self, x32, w32, dy, need_dx=True, need_dw=True, need_db=True):
        b, rp1, k = x32.shape
        n = w32.shape[0]
        dev = x32.device
        work = torch.empty(b * rp1 * n, device=dev, dtype=torch.float32)
        dx = torch.empty((b, rp1, k), device=dev, dtype=torch.float32) if need_dx else None
        dw = torch.empty((n, k), device=dev, dtype=torch.float32) if need_dw else None
        db = torch.empty(n, device=dev, dtype=torch.float32) if need_db else None
        _lib.call("damsm_project_regions_bwd", x32.data_ptr(), b, rp1 - 1, k, w32.data_ptr(), n, dy.data_ptr(),
                  work.data_ptr(), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _stream())
        _lib.add_launches(int(need_dx) + int(need_dw) + int(need_db) - 1)      # two GEMM launches + the column sum
        return dx, dw, db

    # ---- nearest resize in front of the CLIP re-encode (losses.py:348) --------------------------------------------
    def resize_nearest_fwd(self, x, hout, wout):
        """x (..., hin, win) contiguous, 2- or 4-byte elements -> (..., hout, wout)."""
        _require_cuda(x)
        hin, win = x.shape[-2:]
        planes = x.numel() // (hin * win) if hin * win else 0
        y = torch.empty(tuple(x.shape[:-2]) + (hout, wout), device=x.device, dtype=x.dtype)
        _lib.call("damsm_resize_nearest_fwd", x.data_ptr(), x.element_size(), planes, hin, win, hout, wout, y.data_ptr(),
                  _stream())
        return y

    def resize_nearest_bwd(self, dy32, hin, win):
        hout, wout = dy32.shape[-2:]
        planes = dy32.numel() // (hout * wout) if hout * wout else 0
        dx = torch.empty(tuple(dy32.shape[:-2]) + (hin, win), device=dy32.device, dtype=torch.float32)
        _lib.call("damsm_resize_nearest_bwd", dy32.data_ptr(), planes, hin, win, hout, wout, dx.data_ptr(), _stream())
        return dx

    # ---- rm_special_token (pretrain_DAMSM.py:58-79) ------------------------------------------------------------
    def rm_special_token_fwd(self, x, mask_i64):
        """x (B,n,D) with a contiguous innermost dim, 2- or 4-byte elements; mask_i64 (B,n) int64."""
        _require_cuda(x, mask_i64)
        b, n, d = x.shape
        out = torch.empty((b, n - 2, d), device=x.device, dtype=x.dtype)
        out_mask = torch.empty((b, n - 2), device=x.device, dtype=torch.int64)
        _lib.call("damsm_rm_special_token_fwd", x.data_ptr(), x.element_size(), b, n, d, x.stride(0), x.stride(1),
                  mask_i64.data_ptr(), mask_i64.stride(0), mask_i64.stride(1), out.data_ptr(), out_mask.data_ptr(),
                  _stream())
        return out, out_mask

    def rm_special_token_bwd(self, dout, mask_i64, n):
        b, _, d = dout.shape
        dx = torch.empty((b, n, d), device=dout.device, dtype=dout.dtype)
        _lib.call("damsm_rm_special_token_bwd", dout.data_ptr(), dout.element_size(), b, n, d, mask_i64.data_ptr(),
                  mask_i64.stride(0), mask_i64.stride(1), dx.data_ptr(), _stream())
        return dx

    # ---- func_attention -----------------------------------------------------------------------------------------
    def func_attention_fwd(self, qhat, vhat, ctx3, mask_u8, gamma1):
        _require_cuda(qhat, vhat, ctx3, mask_u8)
        b, t, d = qhat.shape
        r = vhat.shape[1]
        dev = qhat.device
        wc = torch.empty((b, t, d), device=dev, dtype=torch.float32)
        attn = torch.empty((b, t, r), device=dev, dtype=torch.float32)
        attn2 = torch.empty((b, t, r), device=dev, dtype=torch.float32)
        csb, csr, csd = ctx3.stride()
        _lib.call("damsm_func_attention_fwd_f32", qhat.data_ptr(), vhat.data_ptr(), ctx3.data_ptr(), csb, csr, csd,
                  mask_u8.data_ptr(), b, t, r, d, float(gamma1), wc.data_ptr(), attn.data_ptr(), attn2.data_ptr(),
                  _stream())
        return wc, attn, attn2

    def func_attention_bwd(self, qhat, vhat, ctx3, attn, attn2, d_wc, d_attn, gamma1):
        b, t, d = qhat.shape
        r = vhat.shape[1]
        dev = qhat.device
        dqhat = torch.empty((b, t, d), device=dev, dtype=torch.float32)
        dvhat = torch.empty((b, r, d), device=dev, dtype=torch.float32)
        dctx = torch.empty((b, r, d), device=dev, dtype=torch.float32)
        csb, csr, csd = ctx3.stride()
        _lib.call("damsm_func_attention_bwd_f32", qhat.data_ptr(), vhat.data_ptr(), ctx3.data_ptr(), csb, csr, csd,
                  attn.data_ptr(), attn2.data_ptr(), _lib.ptr(d_wc), _lib.ptr(d_attn), b, t, r, d, float(gamma1),
                  dqhat.data_ptr(), dvhat.data_ptr(), dctx.data_ptr(), _stream())
        return dqhat, dvhat, dctx


def _dense_strides_like(x: torch.Tensor):
    """Strides of a dense tensor whose dimension order (by stride) matches ``x`` -- the layout autograd
    would pick for a gradient flowing back through the same permute/slice."""
    order = sorted(range(x.dim()), key=lambda i: (x.stride(i), -i), reverse=True)
    strides = [0] * x.dim()
    acc = 1
    for i in reversed(order):
        strides[i] = acc
        acc *= x.shape[i]
    return tuple(strides)


_default_engines = {}


def get_engine(precision: str = "fp32"):
    if precision not in _default_engines:
        _default_engines[precision] = CudaEngine(precision)
    return _default_engines[precision]
