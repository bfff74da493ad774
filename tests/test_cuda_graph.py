"""The C ABI launches asynchronously on the caller's stream and allocates nothing itself, so the exact-fp32 path
(BASELINE configs[1], C2: B=48, T=18, R=49, class mask off) can be captured into a CUDA graph and replayed -- the way
to take the per-launch host overhead out of the latency-bound small configurations."""
import importlib

import numpy as np
import pytest
import torch

from oracle import damsm_oracle as O

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("t2i_clip-gan_b200")
GAM = (4.0, 5.0, 10.0)


def test_c2_step_captured_in_a_cuda_graph_matches_eager_and_oracle():
    B, T, R = 48, 18, 49
    x = O.make_inputs(B, T, R, seed=2027, class_ids=False)
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], None, *GAM)
    os_ = O.sent_loss(x["img"], x["sent"], x["labels"], None, GAM[2])
    dev = "cuda"
    words = torch.tensor(x["words"], device=dev).requires_grad_(True)
    regions = torch.tensor(x["regions"], device=dev).requires_grad_(True)
    img = torch.tensor(x["img"], device=dev).requires_grad_(True)
    sent = torch.tensor(x["sent"], device=dev).requires_grad_(True)
    mask = torch.tensor(x["mask"], device=dev)
    labels = torch.arange(B, device=dev)

    def step():
        w0, w1, _ = pkg.words_loss(regions.permute(0, 2, 1), words.permute(0, 2, 1), labels, None, None, B, mask, *GAM)
        s0, s1 = pkg.sent_loss(img, sent, labels, None, B, gamma3=GAM[2])
        loss = w0 + w1 + s0 + s1
        grads = torch.autograd.grad(loss, (words, regions, img, sent))
        return torch.stack([w0, w1, s0, s1]).detach(), grads

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                      # warm-up on the side stream (allocator pools, lazy init)
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    eager_losses, eager_grads = step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        graph_losses, graph_grads = step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(graph_losses, eager_losses)
    for a, b in zip(graph_grads, eager_grads):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)        # the backward accumulates with fp32 atomics
    ref = np.array([o["loss0"], o["loss1"], os_["loss0"], os_["loss1"]])
    assert np.abs(graph_losses.cpu().numpy() - ref).max() <= 1e-5
    assert np.abs(graph_grads[0].cpu().numpy() - o["dwords"]).max() <= 1e-5 * np.abs(o["dwords"]).max()
    assert np.abs(graph_grads[1].cpu().numpy() - o["dregions"]).max() <= 1e-5 * np.abs(o["dregions"]).max()
    # new inputs through the same graph: refill the static tensors, replay
    x2 = O.make_inputs(B, T, R, seed=99, class_ids=False)
    o2 = O.words_loss(x2["words"], x2["regions"], x["mask"], x["labels"], None, *GAM)
    with torch.no_grad():
        words.copy_(torch.tensor(x2["words"]))
        regions.copy_(torch.tensor(x2["regions"]))
    g.replay()
    torch.cuda.synchronize()
    assert abs(float(graph_losses[0]) - o2["loss0"]) <= 1e-5 and abs(float(graph_losses[1]) - o2["loss1"]) <= 1e-5
