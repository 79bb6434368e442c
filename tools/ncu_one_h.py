"""One device-resident launch of the H kernel (BASELINE config 3) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pydegensac_b200 import _cabi
from pydegensac_b200.scenes import scene_H
P = int(sys.argv[1]) if len(sys.argv) > 1 else 592
N = 5000
b1 = np.empty((P, N, 2)); b2 = np.empty((P, N, 2))
for i in range(P):
    b1[i], b2[i], _ = scene_H(N, 1500, i)
dev = torch.device("cuda:0")
d1 = torch.from_numpy(b1).to(dev); d2 = torch.from_numpy(b2).to(dev)
seeds = torch.arange(P, dtype=torch.int64, device=dev)
H = torch.zeros((P, 9), dtype=torch.float64, device=dev)
mask = torch.zeros((P, N), dtype=torch.uint8, device=dev)
stats = torch.zeros((P, 4), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
_cabi.homography_batch_dev(d1.data_ptr(), d2.data_ptr(), P, N, 2, 3.0, 0.999, 10000, 0, True, 0.0, seeds.data_ptr(),
                           H.data_ptr(), mask.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done", float(mask.sum(1).double().mean()), stats[:4].cpu().numpy().tolist())
