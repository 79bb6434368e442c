"""One device-resident launch of the F kernel for ncu (the host-buffer API overlaps its input feed with the kernel and
falls back to a second launch when a profiler serialises the streams -- profile the device-pointer entry point)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pydegensac_b200 import _cabi
if os.environ.get('DGLIB'): _cabi._LIBPATH = os.path.abspath(os.environ['DGLIB'])
from pydegensac_b200.scenes import batch_F
P = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N = 2000
b1, b2 = batch_F(P)
dev = torch.device("cuda:0")
d1 = torch.from_numpy(b1).to(dev); d2 = torch.from_numpy(b2).to(dev)
seeds = torch.arange(P, dtype=torch.int64, device=dev)
F = torch.zeros((P, 9), dtype=torch.float64, device=dev)
mask = torch.zeros((P, N), dtype=torch.uint8, device=dev)
stats = torch.zeros((P, 4), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
_cabi.fundamental_batch_dev(d1.data_ptr(), d2.data_ptr(), P, N, 2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds.data_ptr(),
                            F.data_ptr(), mask.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done", float(mask.sum(1).double().mean()))
