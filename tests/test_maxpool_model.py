"""CPU model of the 2x2-block max-pool backward (k_maxpool_bwd2 in native/ops/fused_ops.cu): same window / tap arithmetic, checked
against autograd.  The CUDA kernel is the default for even H and W (SHIPYARD_MAXPOOL_BWD2=0 selects the per-pixel kernel); tests/test_zz_gpu_bn_dual.py compares the two on hardware."""
import pytest
import torch
import torch.nn.functional as F


def _fwd_codes(x):
    """3x3 / stride 2 / pad 1 max-pool with arg-max codes ky*3+kx, first maximum wins (k_maxpool_fwd)."""
    n, h, w, c = x.shape
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = torch.full((n, oh, ow, c), float("-inf")); code = torch.zeros((n, oh, ow, c), dtype=torch.int64)
    for oy in range(oh):
        for ox in range(ow):
            for ky in range(3):
                iy = 2 * oy - 1 + ky
                if iy < 0 or iy >= h:
                    continue
                for kx in range(3):
                    ix = 2 * ox - 1 + kx
                    if ix < 0 or ix >= w:
                        continue
                    v = x[:, iy, ix, :]
                    better = v > y[:, oy, ox, :]
                    y[:, oy, ox, :] = torch.where(better, v, y[:, oy, ox, :])
                    code[:, oy, ox, :] = torch.where(better, torch.full_like(code[:, oy, ox, :], ky * 3 + kx), code[:, oy, ox, :])
    return y, code


def _bwd2(dy, code, h, w):
    n, oh, ow, c = dy.shape
    dx = torch.zeros((n, h, w, c))
    for a in range(h // 2):
        for b in range(w // 2):
            for wy in range(2):
                oy = a + wy
                if oy >= oh:
                    continue
                for wx in range(2):
                    ox = b + wx
                    if ox >= ow:
                        continue
                    for i in range(2):
                        ky = 2 * a + i - (2 * oy - 1)
                        if ky < 0 or ky > 2:
                            continue
                        for j in range(2):
                            kx = 2 * b + j - (2 * ox - 1)
                            if kx < 0 or kx > 2:
                                continue
                            hit = code[:, oy, ox, :] == ky * 3 + kx
                            dx[:, 2 * a + i, 2 * b + j, :] += torch.where(hit, dy[:, oy, ox, :], torch.zeros_like(dy[:, oy, ox, :]))
    return dx


@pytest.mark.parametrize("h,w", [(8, 8), (6, 10), (112 // 8, 112 // 8), (2, 4)])
def test_maxpool_bwd2_block_arithmetic_matches_autograd(h, w):
    torch.manual_seed(h * 31 + w)
    x = torch.randn(2, h, w, 3)
    y, code = _fwd_codes(x)
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y.permute(0, 3, 1, 2), yr)
    g = torch.randn_like(yr)
    yr.backward(g)
    dx = _bwd2(g.permute(0, 2, 3, 1), code, h, w)
    assert torch.allclose(dx.permute(0, 3, 1, 2), xr.grad, atol=1e-6)
