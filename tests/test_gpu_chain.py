"""GPU: the GEMM chain kernel (csrc/gemm_chain.cuh) -- a DAG of dependent tcgen05 GEMMs in ONE persistent launch -- against
float64 matmuls: forward MLP chains, two independent branches joined by a contraction (the CTRL logits), a backward chain
with MN-major operands and activation-derivative epilogues, forced tile widths / K-splits, ragged shapes, and repeated
launches of one program (the completion counters must re-arm themselves)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(c, ref):
    return ((c.double() - ref).norm() / (ref.norm() + 1e-30)).item()


def _elu(x):
    return torch.nn.functional.elu(x)


# RLREP_CHAIN_DIST: 0 = split-K partials reduced by the last CTA to arrive, 2 = reduction distributed over the tile's split
# CTAs wherever the tile width allows it (each finalises bn / split columns), 1 (default) = the planner's cost model decides
@pytest.mark.parametrize("dist", [0, 2])
@pytest.mark.parametrize("bn,split_k", [(0, 0), (32, 1), (32, 2), (64, 2), (64, 4), (128, 4), (128, 8), (64, 8), (32, 16)])
@pytest.mark.parametrize("B,S,H,D", [(256, 23, 1024, 2048), (128, 40, 256, 256), (100, 23, 96, 160)])
def test_forward_branches_and_logits(lib, monkeypatch, dist, bn, split_k, B, S, H, D):
    """phi(x) and mu(y) (3 layers each, independent) then logits = phi mu^T: 7 GEMMs, 4 levels."""
    from rlrep_b200 import _lib
    monkeypatch.setenv("RLREP_CHAIN_DIST", str(dist))
    torch.manual_seed(0)
    dev = "cuda"
    Sp = (S + 3) // 4 * 4
    x = torch.zeros(B, Sp, device=dev); x[:, :S] = torch.randn(B, S, device=dev)
    y = torch.zeros(B, Sp, device=dev); y[:, :S] = torch.randn(B, S, device=dev)
    W = {}
    for net in "pm":
        W[net] = [(torch.randn(H, Sp, device=dev) / Sp ** 0.5, torch.randn(H, device=dev)),
                  (torch.randn(H, H, device=dev) / H ** 0.5, torch.randn(H, device=dev)),
                  (torch.randn(D, H, device=dev) / H ** 0.5, torch.randn(D, device=dev))]
        W[net][0][0][:, S:] = 0
    h = {net: [torch.full((B, H), float("nan"), device=dev), torch.full((B, H), float("nan"), device=dev),
               torch.full((B, D), float("nan"), device=dev)] for net in "pm"}
    Bp = (B + 3) // 4 * 4
    logits = torch.full((B, Bp), float("nan"), device=dev)
    specs = []
    for net, inp, last in (("p", x, "none"), ("m", y, "tanh")):
        acts = ("elu", "elu", last)
        src = inp
        for l in range(3):
            specs.append((src, W[net][l][0], h[net][l], dict(epi=_lib.make_epilogue(bias=W[net][l][1], act=acts[l]))))
            src = h[net][l]
    specs.append((h["p"][2], h["m"][2], logits[:, :B], {}))
    ms, levels = _lib.gemm_chain(specs, bn=bn, split_k=split_k, iters=3)
    assert levels == 4
    ref = {}
    for net, inp, last in (("p", x, None), ("m", y, torch.tanh)):
        a = inp.double()
        for l in range(3):
            a = a @ W[net][l][0].double().t() + W[net][l][1].double()
            a = _elu(a) if l < 2 else (last(a) if last else a)
            assert _rel(h[net][l], a) < 2e-3, (net, l, _rel(h[net][l], a))
        ref[net] = a
    want = ref["p"] @ ref["m"].t()
    assert _rel(logits[:, :B], want) < 2e-3, _rel(logits[:, :B], want)
    print(f"chain fwd B={B} H={H} D={D} bn={bn} split={split_k}: {ms * 1e3:.1f} us per launch")


@pytest.mark.parametrize("dist", [0, 2])
@pytest.mark.parametrize("bn,split_k", [(0, 0), (64, 4), (128, 2), (128, 8)])
def test_backward_chain(lib, monkeypatch, dist, bn, split_k):
    """dX2 = (dY W3) * elu'(h2); dW3 = dY^T h2; dX1 = (dX2 W2) * elu'(h1); dW2 = dX2^T h1: MN-major operands, derivative
    epilogues, in-chain read-after-write edges."""
    from rlrep_b200 import _lib
    monkeypatch.setenv("RLREP_CHAIN_DIST", str(dist))
    torch.manual_seed(1)
    dev = "cuda"
    B, H, D = 256, 1024, 2048
    dY = torch.randn(B, D, device=dev)
    h2, h1 = _elu(torch.randn(B, H, device=dev)), _elu(torch.randn(B, H, device=dev))
    W3, W2 = torch.randn(D, H, device=dev) / H ** 0.5, torch.randn(H, H, device=dev) / H ** 0.5
    dX2, dX1 = torch.full((B, H), float("nan"), device=dev), torch.full((B, H), float("nan"), device=dev)
    dW3, dW2 = torch.full((D, H), float("nan"), device=dev), torch.full((H, H), float("nan"), device=dev)
    specs = [
        (dY, W3, dX2, dict(b_mn=True, epi=_lib.make_epilogue(aux=h2, dact="elu_out"))),
        (dY, h2, dW3, dict(a_mn=True, b_mn=True)),
        (dX2, W2, dX1, dict(b_mn=True, epi=_lib.make_epilogue(aux=h1, dact="elu_out"))),
        (dX2, h1, dW2, dict(a_mn=True, b_mn=True)),
    ]
    ms, levels = _lib.gemm_chain(specs, bn=bn, split_k=split_k, iters=5)
    assert levels == 2
    d = lambda h: torch.where(h > 0, torch.ones_like(h), h + 1).double()
    r2 = (dY.double() @ W3.double()) * d(h2)
    r1 = (r2 @ W2.double()) * d(h1)
    assert _rel(dX2, r2) < 2e-3 and _rel(dX1, r1) < 2e-3
    assert _rel(dW3, dY.double().t() @ h2.double()) < 2e-3
    assert _rel(dW2, r2.t() @ h1.double()) < 2e-3
    print(f"chain bwd bn={bn} split={split_k}: {ms * 1e3:.1f} us per launch")


@pytest.mark.parametrize("dist", [0, 2])
def test_chain_is_deterministic_and_rearms(lib, monkeypatch, dist):
    """Split-K partials are summed in split order whoever reduces them (the last CTA to arrive, or each split CTA its own
    columns): repeated launches are bit-identical, and so are the two reduction schemes."""
    from rlrep_b200 import _lib
    monkeypatch.setenv("RLREP_CHAIN_DIST", str(dist))
    torch.manual_seed(2)
    dev = "cuda"
    x = torch.randn(256, 1024, device=dev)
    W1, W2 = torch.randn(1024, 1024, device=dev) / 32, torch.randn(512, 1024, device=dev) / 32
    outs = []
    for rep in range(3):
        h, y = torch.empty(256, 1024, device=dev), torch.empty(256, 512, device=dev)
        _lib.gemm_chain([(x, W1, h, dict(epi=_lib.make_epilogue(act="relu"))), (h, W2, y, {})], bn=64, split_k=4, iters=4)
        outs.append((h.clone(), y.clone()))
    for h, y in outs[1:]:
        assert torch.equal(h, outs[0][0]) and torch.equal(y, outs[0][1])
    _SCHEMES[dist] = outs[0]
    if len(_SCHEMES) == 2:
        assert torch.equal(_SCHEMES[0][0], _SCHEMES[2][0]) and torch.equal(_SCHEMES[0][1], _SCHEMES[2][1])


_SCHEMES = {}
