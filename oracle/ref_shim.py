"""Import the UNMODIFIED reference (``/root/reference``, or its byte-compiled hot-path modules under
``oracle/_ref`` built by ``oracle/build_ref.py`` where the sources do not exist).  TEST INFRASTRUCTURE ONLY.

Used by ``oracle/make_golden.py`` (to record golden vectors) and by
``tests/test_oracle_vs_reference.py`` (live comparison, skipped where
``/root/reference`` does not exist, e.g. on the GPU box).  Nothing under
``/root/reference`` is modified or copied; this file only arranges for
``miscc/losses.py``, ``GlobalAttention.py`` (and ``nt_xent.py`` / ``masks.py``) to import:

  * ``easydict`` is not installed -> an in-memory stand-in module
    (needed by miscc/config.py:6);
  * on a CPU-only host losses.py:127 calls ``.cuda()`` and losses.py:145 calls
    ``.get_device()``; both are neutralised on the Tensor class *only while a
    reference function runs* (context manager), and ``cfg.CUDA`` is set False.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
import warnings

SRC_ROOT = "/root/reference/DMGAN+CLIP/code"
BUILT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _built_ok() -> bool:
    tag = os.path.join(BUILT_ROOT, "PYTHON_TAG")
    return (os.path.isfile(os.path.join(BUILT_ROOT, "miscc", "losses.refbc")) and os.path.isfile(tag)
            and open(tag).read().strip() == sys.implementation.cache_tag)


def sources_available() -> bool:
    """The reference SOURCES are present (build container): needed by the ast-extracting helpers."""
    return os.path.isfile(os.path.join(SRC_ROOT, "miscc", "losses.py"))


def available() -> bool:
    """The reference's hot-path modules can be imported (from source, or from oracle/_ref byte code)."""
    return sources_available() or _built_ok()


REF_ROOT = SRC_ROOT if (sources_available() or not _built_ok()) else BUILT_ROOT


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setattr__(k, v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__


class _ByteCodeFinder:
    """Meta-path finder for the byte-compiled reference modules under oracle/_ref (``<module>.refbc``)."""
    NAMES = {"miscc": "miscc/__init__", "miscc.config": "miscc/config", "miscc.losses": "miscc/losses",
             "GlobalAttention": "GlobalAttention", "nt_xent": "nt_xent", "masks": "masks"}

    def __init__(self, root):
        self.root = root

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        import importlib.util
        rel = self.NAMES.get(name)
        if rel is None:
            return None
        file = os.path.join(self.root, rel + ".refbc")
        if not os.path.isfile(file):
            return None
        loader = importlib.machinery.SourcelessFileLoader(name, file)
        return importlib.util.spec_from_file_location(
            name, file, loader=loader, submodule_search_locations=[os.path.join(self.root, "miscc")] if name == "miscc" else None)


_mods = None


def load():
    """Returns (losses_module, GlobalAttention_module, cfg)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference sources not present at " + REF_ROOT)
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    # the reference's package is called ``miscc`` -- make sure ours is not shadowing it
    for name in [k for k in sys.modules if k == "miscc" or k.startswith("miscc.") or k == "GlobalAttention"]:
        del sys.modules[name]
    finder = _ByteCodeFinder(REF_ROOT) if REF_ROOT == BUILT_ROOT else None
    if finder:
        sys.meta_path.insert(0, finder)
    else:
        sys.path.insert(0, REF_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import GlobalAttention as ga          # noqa: E402
            from miscc import losses              # noqa: E402
            from miscc.config import cfg          # noqa: E402
    finally:
        if finder:
            sys.meta_path.remove(finder)
        else:
            sys.path.remove(REF_ROOT)
    # keep them out of sys.modules so the product's own ``miscc`` drop-in can be imported later
    for name in [k for k in sys.modules if k == "miscc" or k.startswith("miscc.") or k == "GlobalAttention"]:
        del sys.modules[name]
    _mods = (losses, ga, cfg)
    return _mods


@contextlib.contextmanager
def cpu_patches(force_cpu: bool = False):
    """Neutralise the two hard-coded CUDA-isms (losses.py:127, :145) on a CPU-only host, or -- ``force_cpu`` -- on a
    GPU box when the reference is to be timed on the host cores (bench.py's CPU baseline)."""
    import torch
    losses, ga, cfg = load()
    if torch.cuda.is_available() and not force_cpu:
        cfg.CUDA = True
        yield
        return
    old_cuda, old_getdev, old_flag = torch.Tensor.cuda, torch.Tensor.get_device, cfg.CUDA
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.get_device = lambda self: self.device
    cfg.CUDA = False
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.Tensor.get_device, cfg.CUDA = old_cuda, old_getdev, old_flag


def ref_words_loss(words_btd, regions_brd, mask, labels, class_ids, gammas, cap_lens=None):
    """Run the reference words_loss (losses.py:219) on (B,T,D)/(B,R,D) numpy inputs, fp32.
    Returns dict(loss0, loss1, dwords, dregions) with gradients of loss0+loss1."""
    import numpy as np
    import torch
    losses, _, _ = load()
    w = torch.tensor(np.asarray(words_btd), dtype=torch.float32, requires_grad=True)
    r = torch.tensor(np.asarray(regions_brd), dtype=torch.float32, requires_grad=True)
    B = w.shape[0]
    m = torch.tensor(np.asarray(mask), dtype=torch.int64)
    lab = torch.tensor(np.asarray(labels), dtype=torch.int64)
    cl = torch.tensor(np.asarray(cap_lens if cap_lens is not None else np.asarray(mask).sum(1)), dtype=torch.int64)
    with cpu_patches():
        l0, l1, attn = losses.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), lab, cl,
                                         None if class_ids is None else np.asarray(class_ids), B, m,
                                         float(gammas[0]), float(gammas[1]), float(gammas[2]))
        (l0 + l1).backward()
    return dict(loss0=float(l0.detach()), loss1=float(l1.detach()), dwords=w.grad.numpy().copy(), dregions=r.grad.numpy().copy(),
                attn0=attn[0].detach().numpy().copy())


def ref_sent_loss(img, txt, labels, class_ids, gamma3):
    import numpy as np
    import torch
    losses, _, cfg = load()
    a = torch.tensor(np.asarray(img), dtype=torch.float32, requires_grad=True)
    b = torch.tensor(np.asarray(txt), dtype=torch.float32, requires_grad=True)
    lab = torch.tensor(np.asarray(labels), dtype=torch.int64)
    cfg.TRAIN.SMOOTH.GAMMA3 = float(gamma3)
    with cpu_patches():
        l0, l1 = losses.sent_loss(a, b, lab, None if class_ids is None else np.asarray(class_ids), a.shape[0])
        (l0 + l1).backward()
    return dict(loss0=float(l0.detach()), loss1=float(l1.detach()), dimg=a.grad.numpy().copy(), dtxt=b.grad.numpy().copy())


def ref_func_attention(query_btd, context_brd, gamma1, query_mask, d_wc=None):
    import numpy as np
    import torch
    _, ga, _ = load()
    q = torch.tensor(np.asarray(query_btd), dtype=torch.float32, requires_grad=True)
    c = torch.tensor(np.asarray(context_brd), dtype=torch.float32, requires_grad=True)
    m = torch.tensor(np.asarray(query_mask), dtype=torch.int64).unsqueeze(1)
    with cpu_patches(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        wc, attn = ga.func_attention(q.permute(0, 2, 1), c.permute(0, 2, 1), float(gamma1), m)
        out = dict(wc=wc.detach().numpy().copy(), attn=attn.detach().numpy().copy())
        if d_wc is not None:
            wc.backward(torch.tensor(np.asarray(d_wc), dtype=torch.float32))
            out["dquery"] = q.grad.numpy().copy()
            out["dcontext"] = c.grad.numpy().copy()
    return out


def ref_nt_xent(z_i, z_j, temperature):
    """Run the reference NT_Xent (nt_xent.py:4-35) with its own mask builder (masks.py:11-17), fp32.
    Returns dict(loss, dz_i, dz_j).  The two files are loaded under private module names."""
    import importlib.util
    import numpy as np
    import torch
    mods = []
    for name in ("nt_xent", "masks"):
        path = os.path.join(REF_ROOT, name + ".py")
        if os.path.isfile(path):
            spec = importlib.util.spec_from_file_location("_ref_" + name, path)
        else:
            import importlib.machinery
            bc = os.path.join(REF_ROOT, name + ".refbc")
            spec = importlib.util.spec_from_file_location("_ref_" + name, bc,
                                                          loader=importlib.machinery.SourcelessFileLoader("_ref_" + name, bc))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    ntx, masks = mods
    a = torch.tensor(np.asarray(z_i), dtype=torch.float32, requires_grad=True)
    b = torch.tensor(np.asarray(z_j), dtype=torch.float32, requires_grad=True)
    B = a.shape[0]
    crit = ntx.NT_Xent(B, float(temperature), masks.mask_correlated_samples_2(B), torch.device("cpu"))
    loss = crit(a, b)
    loss.backward()
    return dict(loss=float(loss.detach()), dz_i=a.grad.numpy().copy(), dz_j=b.grad.numpy().copy())


def ref_rm_special_token(mask, words_emb):
    """Run the reference's own ``rm_special_token`` (pretrain_DAMSM.py:58-79).  The script cannot be imported (its
    top level needs tensorboardX, nltk, ...), so the function's source is cut out of the file with ``ast`` at run time
    and executed in a namespace that only holds torch -- nothing is copied into this repository."""
    import ast
    import numpy as np
    import torch
    path = os.path.join(SRC_ROOT, "pretrain_DAMSM.py")
    src = open(path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "rm_special_token")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    emb, m = ns["rm_special_token"](torch.as_tensor(np.asarray(mask)), torch.as_tensor(np.asarray(words_emb)))
    return emb.numpy().copy(), m.numpy().copy()


def ref_step(x, gammas):
    """One pass of the reference's own hot path on the HOST cores: words_loss + sent_loss forward + backward
    (losses.py:219-272, :51-91) on torch-CPU tensors, gradients to all four inputs.  ``x``: dict of numpy arrays
    (words (B,T,D), regions (B,R,D), sent, img (B,D), mask (B,T), cap_len, class_ids|None, labels).  What
    ``bench.py --impl reference`` / ``cpu_baseline`` time when the reference is importable (kind "reference")."""
    import numpy as np
    import torch
    losses, _, cfg = load()
    w = torch.tensor(np.asarray(x["words"]), dtype=torch.float32, requires_grad=True)
    r = torch.tensor(np.asarray(x["regions"]), dtype=torch.float32, requires_grad=True)
    a = torch.tensor(np.asarray(x["img"]), dtype=torch.float32, requires_grad=True)
    t = torch.tensor(np.asarray(x["sent"]), dtype=torch.float32, requires_grad=True)
    B = w.shape[0]
    m = torch.tensor(np.asarray(x["mask"]), dtype=torch.int64)
    lab = torch.tensor(np.asarray(x["labels"]), dtype=torch.int64)
    cl = torch.tensor(np.asarray(x["cap_len"]), dtype=torch.int64)
    cls = None if x.get("class_ids") is None else np.asarray(x["class_ids"])
    cfg.TRAIN.SMOOTH.GAMMA3 = float(gammas[2])
    with cpu_patches(force_cpu=True):
        w0, w1, _ = losses.words_loss(r.permute(0, 2, 1), w.permute(0, 2, 1), lab, cl, cls, B, m,
                                      float(gammas[0]), float(gammas[1]), float(gammas[2]))
        s0, s1 = losses.sent_loss(a, t, lab, cls, B)
        (w0 + w1 + s0 + s1).backward()
    return [float(v.detach()) for v in (w0, w1, s0, s1)]


def ref_step_gpu(x, gammas, device="cuda"):
    """The same pass as ``ref_step`` with the UNMODIFIED reference running eagerly on the GPU (its own cfg.CUDA path,
    no patches): the "same-GPU eager PyTorch" baseline of SURVEY 8(d) for the small configurations.  Returns a closure
    that runs one forward + backward (inputs already on the device) and returns the four losses as a tensor."""
    import numpy as np
    import torch
    losses, _, cfg = load()
    cfg.CUDA = True
    cfg.TRAIN.SMOOTH.GAMMA3 = float(gammas[2])
    t = {k: torch.tensor(np.asarray(x[k]), dtype=torch.float32, device=device).requires_grad_(True)
         for k in ("words", "regions", "img", "sent")}
    B = t["words"].shape[0]
    m = torch.tensor(np.asarray(x["mask"]), dtype=torch.int64)            # arrives on the CPU (pretrain_DAMSM.py:110)
    lab = torch.tensor(np.asarray(x["labels"]), dtype=torch.int64, device=device)
    cl = torch.tensor(np.asarray(x["cap_len"]), dtype=torch.int64)
    cls = None if x.get("class_ids") is None else np.asarray(x["class_ids"])

    def step():
        for v in t.values():
            v.grad = None
        w0, w1, _ = losses.words_loss(t["regions"].permute(0, 2, 1), t["words"].permute(0, 2, 1), lab, cl, cls, B, m,
                                      float(gammas[0]), float(gammas[1]), float(gammas[2]))
        s0, s1 = losses.sent_loss(t["img"], t["sent"], lab, cls, B)
        (w0 + w1 + s0 + s1).backward()
        return torch.stack([w0.detach(), w1.detach(), s0.detach(), s1.detach()])
    return step


def ref_r_precision(img_code, sent_codes):
    """Run the reference's own R-precision statements (trainer.py:596-601, inside ``condGANTrainer.sampling``'s
    per-image loop) on ``img_code`` (B, D) and ``sent_codes`` (B, C, D) (true caption at index 0, as :593 builds it).
    The method cannot be imported or called (it needs the GAN, the dataset and CLIP), so the five assignments that
    compute ``scores0`` and the ``torch.argmax(scores0) == 0`` test are cut out of the file with ``ast`` at run time and
    executed per image -- nothing is copied into this repository.  Returns (scores0 (B, C), hit (B,) bool)."""
    import ast
    import numpy as np
    import torch
    path = os.path.join(SRC_ROOT, "trainer.py")
    tree = ast.parse(open(path).read())
    wanted = ["scores", "img_code_norm", "sent_code_norm", "norm", "scores0"]
    found, test = {}, None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "sampling":
            for sub in ast.walk(node):
                if isinstance(sub, ast.Assign) and len(sub.targets) == 1 and isinstance(sub.targets[0], ast.Name) \
                        and sub.targets[0].id in wanted and sub.targets[0].id not in found:
                    found[sub.targets[0].id] = sub
                if isinstance(sub, ast.If) and "argmax" in ast.unparse(sub.test) and test is None:
                    test = sub.test
    assert set(found) == set(wanted) and test is not None, "trainer.py:596-601 not found"
    body = [found[k] for k in wanted] + [ast.Assign(targets=[ast.Name(id="_hit", ctx=ast.Store())], value=test, lineno=0)]
    mod = ast.fix_missing_locations(ast.Module(body=body, type_ignores=[]))
    code = compile(mod, path, "exec")
    img = torch.as_tensor(np.asarray(img_code), dtype=torch.float32)
    cand = torch.as_tensor(np.asarray(sent_codes), dtype=torch.float32)
    scores, hits = [], []
    for i in range(img.shape[0]):
        ns = {"torch": torch, "img_code": img, "sent_code": cand[i], "i": i}
        exec(code, ns)
        scores.append(ns["scores0"].reshape(-1).numpy().copy())
        hits.append(bool(ns["_hit"]))
    return np.stack(scores), np.asarray(hits)
