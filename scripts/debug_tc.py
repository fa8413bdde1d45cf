"""Bring-up aid for the tcgen05 decode_bwd kernel: dumps raw TMEM and compares GEMM2 layouts."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ups_b200
from ups_b200 import _cabi as C

lib = C.lib
fn = lib.ups_debug_decode_bwd_tc
fn.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
fn.restype = ctypes.c_int
B, P, K, F = 1, 128, 16, 64
g = torch.Generator().manual_seed(0)
l0 = torch.randn(B, P, K, generator=g)
m0 = ups_b200.ops.part_softmax(l0.cuda().view(B, P, 1, K)).view(B, P, K).contiguous()
feat = torch.randn(B, K, F, generator=g).cuda()
g_inj = torch.randn(B, P, F + K, generator=g).cuda()
mh = (m0 == m0.max(-1, keepdim=True).values).float()
want = torch.einsum('bpf,bpk->bfk', g_inj[..., :F], mh)[0]     # [F, K]
ws = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
dl0 = torch.zeros(B, P, K, device='cuda'); dfeat = torch.zeros(B, K, F, device='cuda')
dbg = torch.full((128 * 32,), -7.0, device='cuda')
rc = fn(g_inj.data_ptr(), m0.data_ptr(), None, feat.data_ptr(), dl0.data_ptr(), dfeat.data_ptr(), B, P, K, F,
        ws.data_ptr(), ws.numel(), None, dbg.data_ptr(), 0)
torch.cuda.synchronize()
tm = dbg.view(128, 32).cpu()
print(f"rc={rc}: D1 nonzero lanes {(tm[:, :K].abs().sum(1) > 0).sum().item()}")
print("dfeat err", (dfeat[0].t().cpu() - want.cpu()).abs().max().item(), "|want|max", want.abs().max().item())
