"""ncu target: the default 2-CTA tcgen05 GEMM on four GPT-2-XL shapes (one warm-up + one profiled launch each).
   ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 4 -c 4 -o gpurun_out/prof_gemm2 python tools/ncu_gemm.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyparallellibrary_b200.ops import linear as L

T, d = 8192, 1600
x = torch.randn(T, d, device="cuda").bfloat16()
w_qkv = (torch.randn(3 * d, d, device="cuda") * 0.02).bfloat16()
w_fc1 = (torch.randn(4 * d, d, device="cuda") * 0.02).bfloat16()
dy_fc1 = torch.randn(T, 4 * d, device="cuda").bfloat16()
dy_proj = torch.randn(T, d, device="cuda").bfloat16()
h = torch.randn(T, 4 * d, device="cuda").bfloat16()
cases = [lambda: L.gemm(x, w_qkv),                                          # qkv fwd   (NT)
         lambda: L.gemm(dy_fc1, w_fc1, b_mn_major=True),                    # fc1 dX    (NN)
         lambda: L.gemm(dy_proj, h, a_mn_major=True, b_mn_major=True),      # fc2 dW    (TN) [1600 x 6400]
         lambda: L.gemm(dy_proj, x, a_mn_major=True, b_mn_major=True)]      # proj dW   (TN) [1600 x 1600]
for c in cases:      # warm-up: launches 0..3
  c()
torch.cuda.synchronize()
for c in cases:      # profiled: launches 4..7
  c()
torch.cuda.synchronize()
