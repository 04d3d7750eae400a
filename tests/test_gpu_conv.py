"""GPU: the pixel encoder (augmentation + normalisation + 4 convolutions, forward and backward) through the C ABI
against a plain PyTorch fp32 restatement of the reference modules (network_arch/drqv2.py:21-57 RandomShiftsAug with its
grid_sample arithmetic, :138-167 Encoder).  fp32 mode: 1e-5; tf32 mode: 1e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ref_aug(x, shift, pad=4):
    """RandomShiftsAug.forward (drqv2.py:29-57) with the integer draw `shift` [n, 1, 1, 2] injected."""
    n, c, h, w = x.size()
    x = F.pad(x, (pad,) * 4, "replicate")
    eps = 1.0 / (h + 2 * pad)
    arange = torch.linspace(-1.0 + eps, 1.0 - eps, h + 2 * pad, device=x.device, dtype=x.dtype)[:h]
    arange = arange.unsqueeze(0).repeat(h, 1).unsqueeze(2)
    base_grid = torch.cat([arange, arange.transpose(1, 0)], dim=2).unsqueeze(0).repeat(n, 1, 1, 1)
    grid = base_grid + shift.to(x.dtype) * (2.0 / (h + 2 * pad))
    return F.grid_sample(x, grid, padding_mode="zeros", align_corners=False)


def ref_encoder(sd, obs):
    h = obs / 255.0 - 0.5
    h = F.relu(F.conv2d(h, sd["convnet.0.weight"], sd["convnet.0.bias"], stride=2))
    for i in (2, 4, 6):
        h = F.relu(F.conv2d(h, sd[f"convnet.{i}.weight"], sd[f"convnet.{i}.bias"], stride=1))
    return h.flatten(1)


@pytest.mark.parametrize("channels,batch", [(9, 8), (3, 5), (9, 64)])
@pytest.mark.parametrize("regime", ["all_active", "realistic"])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 3e-3)])
def test_conv_encoder_matches_torch(channels, batch, regime, precision, tol):
    """Two weight regimes.
    all_active: non-negative weights behind a +8 bias keep every pre-activation positive, so the ReLU masks are all
      ones on both sides and forward AND backward must agree to the precision's bar (this pins im2col / GEMM / col2im /
      layout permutations exactly).
    realistic: He-scaled signed weights, half the units masked.  The forward pass keeps the bar; the gradient of a
      ReLU network is discontinuous -- a pre-activation within rounding distance of zero flips its mask between two
      correct implementations (measured: ~1 of 2.5e6 per layer in fp32, ~1e-3 of them under TF32 operand rounding) and
      every flip moves the gradient by a whole dfeat entry -- so gradients are held to 5e-3 (fp32) / 6e-2 (tf32),
      which is what PyTorch's own allow_tf32 paths show against fp32."""
    from rlrep_b200.pixel import ConvEncoder
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(channels * 100 + batch)
    sd = {}
    for i, cin in zip((0, 2, 4, 6), (channels, 32, 32, 32)):
        w = torch.randn(32, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
        b = torch.randn(32, generator=g) * 0.1
        if regime == "all_active":
            w, b = (w.abs() / 8 if i > 0 else w), b + 8.0
        sd[f"convnet.{i}.weight"], sd[f"convnet.{i}.bias"] = w, b
    obs = torch.randint(0, 256, (batch, channels, 84, 84), generator=g, dtype=torch.uint8)
    shift = torch.randint(0, 9, (batch, 1, 1, 2), generator=g)
    dfeat = torch.randn(batch, 32 * 35 * 35, generator=g)

    enc = ConvEncoder((channels, 84, 84), batch=batch, precision=precision)
    enc.load_state_dict(sd)
    for k, v in enc.state_dict().items():
        assert torch.equal(v, sd[k]), k  # layout round trip through the C ABI
    feat = enc.forward(obs.cuda(), shift.reshape(batch, 2).cuda())
    enc.backward(dfeat.cuda())
    grads = enc.grads()
    feat_plain = enc.forward(obs.cuda(), None)

    rsd = {k: v.cuda().requires_grad_() for k, v in sd.items()}
    ref = ref_encoder(rsd, ref_aug(obs.cuda().float(), shift.cuda()))
    ref.backward(dfeat.cuda())
    ref_plain = ref_encoder({k: v.detach() for k, v in rsd.items()}, obs.cuda().float())

    def rel(a, b):
        return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()

    errs = {"feat": rel(feat, ref.detach()), "feat_noaug": rel(feat_plain, ref_plain)}
    for k in sd:
        errs[k] = rel(grads[k].cuda(), rsd[k].grad)
    print(f"C={channels} B={batch} {regime} {precision}: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert errs["feat"] < tol and errs["feat_noaug"] < tol, errs
    grad_tol = tol if regime == "all_active" else (5e-3 if precision == "fp32" else 6e-2)
    assert max(v for k, v in errs.items() if k.startswith("convnet")) < grad_tol, errs
    enc.close()
