"""Phase timeline of the attention forward kernel (CTA 0,0,0; clock64 stamps written by the kernel when a debug buffer is set)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easyparallellibrary_b200.ops import _lib
from easyparallellibrary_b200.ops.attention_kernel import flash_attention_packed

lib = _lib.require()
lib.epl_attn_set_debug.argtypes = [ctypes.c_void_p]
lib.epl_attn_set_debug.restype = None
B, S, H, D = 8, 1024, 25, 64
qkv = (torch.randn(B, S, 3, H, D, device="cuda") * 0.5).bfloat16()
for _ in range(3):
  flash_attention_packed(qkv, True)
dbg = torch.zeros(128 * 8, dtype=torch.int64, device="cuda")
lib.epl_attn_set_debug(dbg.data_ptr())
flash_attention_packed(qkv, True)
torch.cuda.synchronize()
lib.epl_attn_set_debug(None)
t = dbg.cpu().view(128, 8)
t0 = int(t[64, 0])
print("softmax warp 2 (cycles): wait_S | tmem_ld | max+xchg | (rescale) | exp+P store | fence+arrive || step total")
for j in range(8):
  r = [int(x) for x in t[j, :7]]
  print("  j=%d  start %6d | %5d | %5d | %5d | %5d | %5d | %5d || %6d" % (j, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[6] - r[0]))
print("MMA thread: wait_KV | issue S | wait_P | issue PV")
for j in range(8):
  r = [int(x) for x in t[64 + j, :5]]
  print("  j=%d  start %6d | %5d | %5d | %5d | %5d" % (j, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3]))
