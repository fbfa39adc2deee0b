"""Developer aid: dump the per-phase clock64 timeline of one CTA of attn_bwd_tc_kernel (needs a library built with
AVS_EXTRA_NVCC_FLAGS=-DAVS_TC_TRACE)."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops, _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
import sys as _s
n_seq, S, H, hd = (64, 708, 16, 32) if len(_s.argv) < 2 else (256, int(_s.argv[1]), 12, 64)
D = H * hd
qkv = torch.randn(n_seq * S, 3 * D, device="cuda").bfloat16()
out = torch.empty(n_seq * S, D, device="cuda", dtype=torch.bfloat16)
dout = torch.randn_like(out)
lse = torch.empty(n_seq, H, S, device="cuda"); delta = torch.empty_like(lse); dqkv = torch.empty_like(qkv)
ops.attention_fwd(qkv, out, lse, n_seq, S, H, hd)
trace = torch.zeros(20 * 512, dtype=torch.int64, device="cuda")
ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, n_seq, S, H, hd)
lib.avs_debug_set_tc_trace.argtypes = [ctypes.c_void_p]
lib.avs_debug_set_tc_trace(ctypes.c_void_p(trace.data_ptr()))
ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, n_seq, S, H, hd)
torch.cuda.synchronize()
tr = trace.cpu().reshape(20, 512)
nz = tr[tr > 0]
t0 = int(nz.min())
def rel(k, i):
    v = int(tr[k, i]); return v - t0 if v else None
print("MMA thread per sub-step u: [before wait p_full, got p_full, issued dV/dK(/dQ)+dP^T]")
for u in range(0, 26):
    print(u, [rel(k, u) for k in (1, 2, 12)])
print("softmax group A | B per block-step n: [top, got S^T, got dP^T, arrived p_full]")
for n in range(0, 13):
    print(n, [rel(k, n) for k in (3, 4, 5, 6)], "|", [rel(k, n) for k in (7, 8, 9, 10)])
print("CTA lifetime stamps [entry, after prologue sync, a softmax warp done, after final sync]:", [rel(17, i) for i in range(4)])
print("total cycles", int(tr.max()) - t0)
