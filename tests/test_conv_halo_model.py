"""CPU model of the halo-load 3x3 convolution kernel's index arithmetic (native/gemm/conv_halo.inc).

The kernel itself needs a B200 (tests/test_gpu_gemm.py); what CAN be checked without one is the geometry it relies on: one
zero-filled box of (R+2) x (W+2) pixels per tile, every filter tap = the same array shifted by r*(W+2)+s rows, junk rows
(padding columns, tail) never written, valid rows never reading a row outside the box, and the dgrad variant's reversed
tap / transposed weight read.  The model below does exactly what the producer / MMA / epilogue warps do, in fp32.
"""
import pytest
import torch
import torch.nn.functional as F

from batch_shipyard_b200.ops import gemm


def _halo_model(act: torch.Tensor, w: torch.Tensor, dgrad: bool) -> torch.Tensor:
    """act [N,H,W,Ca], w [Co,3,3,Ci] (KRSC) -> out [N,H,W,Cn]"""
    n, h, wd, ca = act.shape
    cn = w.shape[3] if dgrad else w.shape[0]
    rows = gemm.halo_rows(h, wd)
    assert rows > 0
    wp = wd + 2
    out = torch.full((n, h, wd, cn), float("nan"))
    written = torch.zeros((n, h, wd), dtype=torch.int32)
    for img in range(n):
        for h0 in range(0, h, rows):
            # TMA box {C, Wp, R+2, 1} at (w, h) = (-1, h0 - 1): out-of-bounds elements are zero
            box = torch.zeros((rows + 2, wp, ca))
            for bh in range(rows + 2):
                for bw in range(wp):
                    hh, ww = h0 - 1 + bh, -1 + bw
                    if 0 <= hh < h and 0 <= ww < wd:
                        box[bh, bw] = act[img, hh, ww]
            smem = torch.full((256, ca), float("nan"))            # the slot is 256 rows; rows past the box are stale
            smem[: (rows + 2) * wp] = box.reshape(-1, ca)
            acc = torch.zeros((128, cn))
            for tap in range(9):
                r, s = divmod(tap, 3)
                a = smem[r * wp + s: r * wp + s + 128]            # descriptor start shifted by r*Wp + s rows
                if dgrad:
                    b = w[:, 2 - r, 2 - s, :]                     # W[co][8 - tap][ci] as B[k = co, n = ci]
                    acc = acc + torch.nan_to_num(a, nan=float("nan")) @ b
                else:
                    b = w[:, r, s, :]                             # W[co][tap][ci] as B[n = co, k = ci]
                    acc = acc + a @ b.t()
            for m in range(128):
                hl, wl = divmod(m, wp)
                if hl < rows and wl < wd:
                    out[img, h0 + hl, wl] = acc[m]
                    written[img, h0 + hl, wl] += 1
    assert int(written.min()) == 1 and int(written.max()) == 1   # every output pixel exactly once
    return out


@pytest.mark.parametrize("h,wd", [(7, 7), (14, 14), (28, 28), (56, 56), (6, 10), (16, 16), (4, 4), (8, 8), (32, 32), (2, 2)])
@pytest.mark.parametrize("dgrad", [False, True])
def test_halo_geometry_matches_conv(h, wd, dgrad):
    torch.manual_seed(h * 10 + dgrad)
    n, ci, co = 2, 8, 12
    x = torch.randn(n, ci, h, wd, requires_grad=True)
    w = torch.randn(co, ci, 3, 3) * 0.2
    y = F.conv2d(x, w, padding=1)
    w_krsc = w.permute(0, 2, 3, 1).contiguous()
    if not dgrad:
        got = _halo_model(x.detach().permute(0, 2, 3, 1).contiguous(), w_krsc, False)
        ref = y.detach().permute(0, 2, 3, 1)
    else:
        dy = torch.randn_like(y)
        (dx,) = torch.autograd.grad(y, x, dy)
        got = _halo_model(dy.permute(0, 2, 3, 1).contiguous(), w_krsc, True)
        ref = dx.permute(0, 2, 3, 1)
    assert not torch.isnan(got).any()                              # no valid row ever touched a stale shared-memory row
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)


def test_halo_rows_and_shape_filter():
    assert [gemm.halo_rows(s, s) for s in (7, 14, 28, 56)] == [7, 7, 4, 2]
    assert gemm.halo_rows(112, 112) == 0                          # one padded row (114) still fits, but (R+2)*Wp > 256
    for hh, ww in [(7, 7), (14, 14), (28, 28), (56, 56), (32, 32), (20, 60)]:
        r = gemm.halo_rows(hh, ww)
        if r:
            assert hh % r == 0 and r * (ww + 2) <= 128 and (r + 2) * (ww + 2) <= 256
            assert 127 + 2 * (ww + 2) + 2 < 256                   # the last MMA row of the last tap stays inside the slot
    assert gemm.halo_ok(256, 56, 56, 64, 64, 3, 3, 1, 1)
    assert not gemm.halo_ok(256, 56, 56, 64, 64, 3, 3, 2, 1)      # stride 2: im2col kernels / cuDNN
    assert not gemm.halo_ok(256, 56, 56, 64, 64, 1, 1, 1, 0)
    assert not gemm.halo_ok(256, 56, 56, 32, 64, 3, 3, 1, 1)      # channel blocks of 64
    assert gemm.halo_ok(256, 28, 28, 128, 128, 3, 3, 1, 1, pair=True)
    assert not gemm.halo_ok(256, 56, 56, 64, 64, 3, 3, 1, 1, pair=True)   # CTA pairs need out channels % 128
    assert not gemm.halo_ok(1, 7, 7, 512, 512, 3, 3, 1, 1, pair=True)     # odd number of M tiles


def _wgrad_halo_model(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """x [N,H,W,Ci], dy [N,H,W,Co] -> dW [Co,3,3,Ci]; mirrors native/gemm/wgrad_halo.inc (one X halo box + one dY box in padded
    row coordinates per tile, five tap-pair accumulators of 128 rows = (tap parity, ci), k-steps of 16 pixel rows)."""
    n, h, wd, ci = x.shape
    co = dy.shape[3]
    rows = gemm.halo_rows(h, wd)
    wp = wd + 2
    a_rows, b_rows = (rows + 2) * wp, rows * wp
    ksteps = (b_rows + 15) // 16
    acc = torch.zeros((5, 2 * ci, co))                              # [tap pair][slab * ci + c][co]
    for img in range(n):
        for h0 in range(0, h, rows):
            a = torch.zeros((256, ci))                               # rows past the box are zeroed once at kernel start
            box = torch.zeros((rows + 2, wp, ci))
            for bh in range(rows + 2):
                for bw in range(wp):
                    hh, ww = h0 - 1 + bh, bw - 1
                    if 0 <= hh < h and 0 <= ww < wd:
                        box[bh, bw] = x[img, hh, ww]
            a[:a_rows] = box.reshape(-1, ci)
            b = torch.zeros((128, co))
            bbox = torch.zeros((rows, wp, co))                       # TMA box {co, Wp, R} at w = 0: columns W, W+1 are out of bounds -> 0
            bbox[:, :wd] = dy[img, h0:h0 + rows]
            b[:b_rows] = bbox.reshape(-1, co)
            k = ksteps * 16
            for p in range(5):
                ta, tb = 2 * p, 2 * p + 1
                sh_a = (ta // 3) * wp + ta % 3
                sh_b = (tb // 3) * wp + tb % 3 if tb < 9 else sh_a + 1
                assert sh_b + k <= 256 and sh_b > sh_a
                slab = torch.cat([a[sh_a:sh_a + k], a[sh_b:sh_b + k]], dim=1)      # [k, 2*ci]: M index = slab * ci + c
                acc[p] += slab.t() @ b[:k]
    dw = torch.zeros((co, 9, ci))
    for p in range(5):
        for half in range(2):
            tap = 2 * p + half
            if tap < 9:                                              # the dummy partner of tap 8 is discarded
                dw[:, tap, :] = acc[p, half * ci:(half + 1) * ci, :].t()
    return dw.reshape(co, 3, 3, ci)


@pytest.mark.parametrize("h,wd", [(7, 7), (14, 14), (28, 28), (56, 56), (12, 20)])
def test_halo_wgrad_geometry_matches_conv(h, wd):
    torch.manual_seed(h + wd)
    n, ci, co = 2, 8, 6
    x = torch.randn(n, ci, h, wd)
    dy = torch.randn(n, co, h, wd)
    ref = torch.nn.grad.conv2d_weight(x, (co, ci, 3, 3), dy, padding=1)          # [co, ci, 3, 3]
    got = _wgrad_halo_model(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous())
    assert torch.allclose(got.permute(0, 3, 1, 2), ref, atol=1e-3, rtol=1e-4)
