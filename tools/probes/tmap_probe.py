"""Does cuTensorMapEncodeTiled accept the 3-D view (byte in chunk, chunk in channel, channel) of a channel-major uint8
buffer whose row stride is not a multiple of the 1024-byte chunk?  (strides 1024, 72000)"""
import torch
from cuda.bindings import driver as drv
torch.zeros(1, device="cuda")
C, stride = 1000, 72000
buf = torch.zeros(C * stride + 4096, dtype=torch.uint8, device="cuda")
for st in (72000, 72704):
    for box_h in (32, 1):
        for sw in (drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B, drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_NONE):
            r = drv.cuTensorMapEncodeTiled(
                drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf.data_ptr(),
                [drv.cuuint64_t(1024), drv.cuuint64_t((st + 1023) // 1024), drv.cuuint64_t(C)],
                [drv.cuuint64_t(1024), drv.cuuint64_t(st)],
                [drv.cuuint32_t(128), drv.cuuint32_t(box_h), drv.cuuint32_t(1)],
                [drv.cuuint32_t(1), drv.cuuint32_t(1), drv.cuuint32_t(1)],
                drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_NONE,
                drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
            print("stride", st, "box_h", box_h, sw, "->", r[0])
