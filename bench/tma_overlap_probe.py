"""Hardware/driver check for the space-to-depth stem kernel: does cuTensorMapEncodeTiled accept a dimension whose stride (32 B = one
s2d pixel of 16 bf16 channels) is SMALLER than the extent of the faster dimension (64 elements = 4 pixels = one filter row)?
Overlapping strides would present the im2col rows [output pixel][4 px x 16 ch] of the dense 4x4 stem convolution directly to TMA."""
import json
import sys

import torch
from cuda.bindings import driver as drv


def main():
    torch.cuda.init()
    n, hp, wp, c = 4, 115, 115, 16
    x = torch.arange(n * hp * wp * c, device="cuda", dtype=torch.float32).remainder(251).to(torch.bfloat16)
    u64, u32 = drv.cuuint64_t, drv.cuuint32_t
    out = {}
    for name, dims, strides, box in [
        # {64 contiguous elements, W_out = 112 (stride 16 el = 32 B), H rows (stride Wp*16*2 B), N}
        ("overlap_4d", [64, 112, hp, n], [32, wp * c * 2, hp * wp * c * 2], [64, 112, 1, 1]),
        ("overlap_4d_box128", [64, 112, hp, n], [32, wp * c * 2, hp * wp * c * 2], [64, 128, 1, 1]),
        ("plain_4d", [16, wp, hp, n], [c * 2, wp * c * 2, hp * wp * c * 2], [16, 112, 1, 1]),
    ]:
        res = drv.cuTensorMapEncodeTiled(
            drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, len(dims), x.data_ptr(),
            [u64(d) for d in dims], [u64(s) for s in strides], [u32(b) for b in box], [u32(1)] * len(dims),
            drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B,
            drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_128B, drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        out[name] = str(res[0])
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
