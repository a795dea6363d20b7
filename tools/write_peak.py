"""Write-only HBM bandwidth of the device for the output patterns of the posterior kernels."""
import sys, ctypes as C
sys.path.insert(0, '.')
import baynorm_b200 as bn
ctx = bn.Context(0)
for gb in (8,):
    for planes in (1, 20):
        for run in (1, 2, 4, 8, 16, 64):
            v = C.c_double()
            ctx.check(ctx.lib.bn_debug_write_peak(ctx.h, gb << 30, planes, run, C.byref(v)))
            print("write-only %d GiB planes=%d run=%d x 256 B: %.0f GB/s" % (gb, planes, run, v.value))
