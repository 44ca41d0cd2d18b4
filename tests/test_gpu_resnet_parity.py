"""Full-model correctness of the flagship: the repo's ResNet-50 (bf16, NHWC, tcgen05 / fused-BN kernels, dispatcher in every
mode) against torchvision's resnet50 in fp32 with the SAME weights — logits, loss and the gradient of every parameter tensor —
and the fused all-reduce + SGD trainer against a plain fp32 SGD on the full model (parameters compared, not just the loss)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

# Tolerances are CALIBRATED, not guessed: the same fp32 torchvision model is also run under stock bf16 autocast (cuDNN / cuBLAS,
# channels_last), and its per-tensor error against fp32 is the noise floor of bf16 training arithmetic on this network.  The repo's
# run must stay within NOISE_FACTOR x that floor (+ a small absolute slack) on every tensor: a wrong kernel shows up as a tensor
# that is far outside the floor, while honest bf16 rounding cannot fail the test.
NOISE_FACTOR = float(os.environ.get("SHIPYARD_PARITY_FACTOR", "2.0"))
ABS_SLACK = float(os.environ.get("SHIPYARD_PARITY_SLACK", "0.01"))


def _copy_weights(ours, tv):
    """ours (ConvBN.weight / gamma / beta, fc_weight / fc_bias) -> torchvision module (conv / bn / fc), as fp32."""
    def cb(src, conv, bn):
        conv.weight.data.copy_(src.weight.data.float())
        bn.weight.data.copy_(src.gamma.data.float()); bn.bias.data.copy_(src.beta.data.float())
    cb(ours.stem, tv.conv1, tv.bn1)
    tv_blocks = [b for layer in (tv.layer1, tv.layer2, tv.layer3, tv.layer4) for b in layer]
    assert len(tv_blocks) == len(ours.blocks)
    for ob, tb in zip(ours.blocks, tv_blocks):
        cb(ob.c1, tb.conv1, tb.bn1); cb(ob.c2, tb.conv2, tb.bn2); cb(ob.c3, tb.conv3, tb.bn3)
        assert (ob.down is None) == (tb.downsample is None)
        if ob.down is not None:
            cb(ob.down, tb.downsample[0], tb.downsample[1])
    tv.fc.weight.data.copy_(ours.fc_weight.data.float()); tv.fc.bias.data.copy_(ours.fc_bias.data.float())


def _named_pairs(ours, tv):
    """(our parameter name, our parameter, torchvision parameter) for every parameter tensor."""
    out = []

    def cb(name, src, conv, bn):
        out.append((name + ".weight", src.weight, conv.weight))
        out.append((name + ".gamma", src.gamma, bn.weight))
        out.append((name + ".beta", src.beta, bn.bias))
    cb("stem", ours.stem, tv.conv1, tv.bn1)
    tv_blocks = [b for layer in (tv.layer1, tv.layer2, tv.layer3, tv.layer4) for b in layer]
    for i, (ob, tb) in enumerate(zip(ours.blocks, tv_blocks)):
        cb(f"blocks.{i}.c1", ob.c1, tb.conv1, tb.bn1); cb(f"blocks.{i}.c2", ob.c2, tb.conv2, tb.bn2); cb(f"blocks.{i}.c3", ob.c3, tb.conv3, tb.bn3)
        if ob.down is not None:
            cb(f"blocks.{i}.down", ob.down, tb.downsample[0], tb.downsample[1])
    out.append(("fc_weight", ours.fc_weight, tv.fc.weight))
    out.append(("fc_bias", ours.fc_bias, tv.fc.bias))
    return out


def _build(seed=0):
    import torchvision
    from batch_shipyard_b200.models import resnet
    torch.manual_seed(seed)
    ours = resnet.resnet50().cuda().train()
    for m in ours.modules():
        if isinstance(m, resnet.ConvBN):
            m.gamma.data.uniform_(0.5, 1.5)
            m.beta.data.uniform_(-0.2, 0.2)
    for b in ours.blocks:
        # the zero-initialised last gamma of each block would hide the c1/c2/c3 gradients; a small one (as in a trained network)
        # keeps the 16-block residual stack well conditioned
        b.c3.gamma.data.uniform_(0.1, 0.4)
    for p in ours.parameters():
        p.data = p.data.to(torch.bfloat16)
    tv = torchvision.models.resnet50(weights=None).cuda().train().float()
    _copy_weights(ours, tv)
    return ours, tv


def _inputs(n, seed=1):
    from batch_shipyard_b200.ops import fused
    g = torch.Generator(device="cuda").manual_seed(seed)
    img = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g)
    y = torch.randint(0, 1000, (n,), device="cuda", generator=g)
    mean = torch.tensor(fused.IMAGENET_MEAN, device="cuda"); std = torch.tensor(fused.IMAGENET_STD, device="cuda")
    x_ref = ((img.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()
    s2d = torch.empty(n, 224 // 2 + 3, 224 // 2 + 3, 16, dtype=torch.bfloat16, device="cuda")
    fused.u8_to_s2d_norm(img, s2d)
    return img, y, x_ref, s2d.permute(0, 3, 1, 2)


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def _noise_floor(tv, x_ref, y, ref_logits, ref_grads):
    """Per-tensor error of the STOCK bf16 path (autocast, channels_last) against fp32 on the same model and batch."""
    import torch.nn.functional as F
    tv.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lo = tv(x_ref.contiguous(memory_format=torch.channels_last))
    F.cross_entropy(lo.float(), y).backward()
    floor = {name: _rel(p.grad, ref_grads[name]) for name, p in tv.named_parameters()}
    return _rel(lo, ref_logits), floor


def _compare(ours, tv, x_ours, x_ref, y, label):
    import torch.nn.functional as F
    ours.zero_grad(set_to_none=True); tv.zero_grad(set_to_none=True)
    lr = tv(x_ref)
    loss_r = F.cross_entropy(lr, y); loss_r.backward()
    ref_grads = {name: p.grad.detach().clone() for name, p in tv.named_parameters()}
    tv_name = {id(p): name for name, p in tv.named_parameters()}
    lr = lr.detach()
    floor_logits, floor = _noise_floor(tv, x_ref, y, lr, ref_grads)
    lo = ours(x_ours).float()
    loss_o = F.cross_entropy(lo, y); loss_o.backward()
    rel_logits = _rel(lo.detach(), lr)
    rows, bad = [], []
    for name, p_ours, p_tv in _named_pairs(ours, tv):
        go, gr = p_ours.grad, ref_grads[tv_name[id(p_tv)]]
        assert go is not None, (label, name)
        rel = _rel(go, gr)
        fl = floor[tv_name[id(p_tv)]]
        rows.append((name, rel, fl))
        if rel > NOISE_FACTOR * fl + ABS_SLACK:
            bad.append((name, round(rel, 4), round(fl, 4)))
    worst = max(rows, key=lambda r: r[1])
    ratio = max(rows, key=lambda r: r[1] / max(r[2], 1e-6))
    print(f"[{label}] logits rel {rel_logits:.4f} (stock bf16 floor {floor_logits:.4f}); worst gradient {worst[0]} rel {worst[1]:.4f} "
          f"(floor {worst[2]:.4f}); worst vs floor {ratio[0]} {ratio[1]:.4f}/{ratio[2]:.4f}; {len(rows)} tensors; "
          f"loss {float(loss_o):.4f} vs {float(loss_r):.4f}")
    assert rel_logits <= NOISE_FACTOR * floor_logits + ABS_SLACK, (label, "logits", rel_logits, floor_logits)
    assert abs(float(loss_o) - float(loss_r)) < 0.02 * max(1.0, abs(float(loss_r))), (label, float(loss_o), float(loss_r))
    assert not bad, (label, "gradients outside the bf16 noise floor", bad[:8])


@pytest.mark.parametrize("mode", ["auto", "tc", "cudnn"])
def test_resnet50_gradients_match_torchvision_fp32(mode):
    """Per-layer gradient parity with every convolution pass routed by the dispatcher (`auto`), forced onto the repo's tcgen05 kernels
    (`tc`) and forced onto cuDNN (`cudnn`: isolates the fused BN / pool / FC kernels)."""
    from batch_shipyard_b200.ops import conv
    ours, tv = _build()
    _, y, x_ref, x_s2d = _inputs(32)
    conv.set_mode(mode)
    try:
        _compare(ours, tv, x_s2d, x_ref, y, mode)
    finally:
        conv.set_mode("auto")


def test_resnet50_residual_gradient_fusion_matches_torchvision():
    """SHIPYARD_BN_DUAL (two-handle block outputs: the residual-gradient add happens inside the BN backward kernels)."""
    from batch_shipyard_b200.models import resnet
    ours, tv = _build(seed=2)
    _, y, x_ref, x_s2d = _inputs(16, seed=3)
    resnet.set_bn_dual(True)
    try:
        _compare(ours, tv, x_s2d, x_ref, y, "bn_dual")
    finally:
        resnet.set_bn_dual(False)


def test_full_model_trainer_tracks_fp32_sgd_parameters():
    """Three steps of the fused trainer (flat bf16 parameters, fp32 master weights, fused all-reduce + SGD kernel; eager so that every
    step can be compared — the captured-graph step is covered by tests/_trainer_worker.py) on the full ResNet-50 against
    torch.optim.SGD on the fp32 torchvision model.  What is compared is the UPDATE of the fp32 master weights (master - initial) of
    every parameter tensor against the reference update: that isolates the optimiser semantics (1/N scale, weight decay, momentum,
    learning rate) and the gradients from the 2^-9 rounding of the bf16 parameter image."""
    import torch.nn.functional as F
    from batch_shipyard_b200.ops.coll import Communicator
    from batch_shipyard_b200.parallel.ddp import FusedDataParallelTrainer
    ours, tv = _build(seed=4)
    img, y, x_ref, _ = _inputs(32, seed=5)
    pairs = _named_pairs(ours, tv)
    init = {name: pt.detach().clone() for name, _, pt in pairs}
    comm = Communicator(0, 1, device=0, heap_bytes=1 << 30)
    lr, mom, wd = 0.01, 0.9, 1e-4
    tr = FusedDataParallelTrainer(ours, comm, (32, 3, 224, 224), 1000, lr=lr, momentum=mom, weight_decay=wd, use_graph=False)
    tr.load_images_u8(img, y)
    opt = torch.optim.SGD(tv.parameters(), lr=lr, momentum=mom, weight_decay=wd)
    lay = tr.flat.layout.offsets
    for step in range(3):
        loss = float(tr.step())
        opt.zero_grad(); lref = F.cross_entropy(tv(x_ref), y); lref.backward(); opt.step()
        assert abs(loss - float(lref)) < 0.03 * max(1.0, abs(float(lref))), (step, loss, float(lref))
        worst, num, den = ("", 0.0), 0.0, 0.0
        for name, p_ours, p_tv in pairs:
            off, cnt, shape, cl = lay[name]
            m = tr.flat.master[off:off + cnt]
            if cl:
                co, ci, kh, kw = shape
                m = m.view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                m = m.view(shape)
            upd, ref_upd = m - init[name], p_tv.detach() - init[name]
            rel = float((upd - ref_upd).norm() / ref_upd.norm().clamp_min(1e-12))
            num += float((upd - ref_upd).norm()) ** 2; den += float(ref_upd.norm()) ** 2
            if rel > worst[1]:
                worst = (name, rel)
            # single tensors of this random-init network have a bf16 gradient noise floor of up to ~0.6 (measured against stock
            # bf16 autocast in the gradient test above): per tensor only sign and magnitude are asserted ...
            assert rel < 2.0, (step, name, rel)
            # the bf16 parameter image the kernels read is the rounded master weight
            assert float((p_ours.detach().float() - m).abs().max()) <= float(m.abs().max()) * 2 ** -8, (step, name)
        total = (num / max(den, 1e-30)) ** 0.5
        print(f"[trainer step {step}] loss {loss:.4f} vs {float(lref):.4f}; whole-model update rel {total:.4f}; worst tensor {worst[0]} rel {worst[1]:.4f}")
        # ... and the update of the WHOLE model (dominated by the well-conditioned tensors) must match: the exact optimiser
        # arithmetic (1/N, weight decay, momentum, lr, master/bf16 image, gradient zeroing) is checked against fp64 in
        # tests/_coll_worker.py (fused_sgd), world 1..8
        assert total < 0.6, (step, total, worst)             # (measured 0.33 at step 0: the bf16 gradient noise of this random-init network)
    comm.check_status()
    assert float(tr.flat.grads.abs().max()) == 0.0            # the fused kernel left the gradient buffer zeroed
    comm.close()
