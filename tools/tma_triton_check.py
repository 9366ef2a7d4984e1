"""development diagnostic: does a tensor-descriptor (TMA) load work on this box at all?"""
import torch, triton, triton.language as tl
from triton.tools.tensor_descriptor import TensorDescriptor

@triton.jit
def k(desc, out_ptr, BM: tl.constexpr, BN: tl.constexpr):
    t = desc.load([1, 2, 3])
    t = tl.reshape(t, [BM, BN])
    offs = tl.arange(0, BM)[:, None] * BN + tl.arange(0, BN)[None, :]
    tl.store(out_ptr + offs, t)

a = torch.arange(4 * 20 * 64, device='cuda', dtype=torch.float64).reshape(4, 20, 64)
out = torch.empty(8 * 32, device='cuda', dtype=torch.float64)
d = TensorDescriptor(a, a.shape, a.stride(), [1, 8, 32])
h = k[(1,)](d, out, 8, 32)
torch.cuda.synchronize()
print('ok', torch.equal(out.reshape(8, 32), a[1, 2:10, 3:35]))
ptx = h.asm['ptx']
open('gpurun_out/triton_tma.ptx', 'w').write(ptx)
open('gpurun_out/triton_tma.cubin', 'wb').write(h.asm['cubin'])
import triton.backends.nvidia.driver as drv, inspect, os
src = os.path.join(os.path.dirname(drv.__file__), 'driver.c')
print(src, os.path.exists(src))
txt = open(src).read() if os.path.exists(src) else inspect.getsource(drv)
i = txt.find('cuTensorMapEncodeTiled')
while i >= 0 and i < len(txt):
    print(txt[max(0, i - 1500):i + 700]); print('=' * 80)
    i = txt.find('cuTensorMapEncodeTiled', i + 2000)
    break
