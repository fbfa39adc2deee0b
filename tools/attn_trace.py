"""Developer aid: dump the per-phase clock64 timeline of one CTA of attn_bwd_tc_kernel (needs a library built with
AVS_EXTRA_NVCC_FLAGS=-DAVS_TC_TRACE)."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops, _lib
lib = ctypes.CDLL(os.path.join(os.path.dirname(_lib.__file__), "libavsiam_b200.so"))
n_seq, S, H, hd = 64, 708, 16, 32
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
t0 = int(tr[0, 0])
names = ["mma:issueS", "mma:wait_p", "mma:got_p", "A:top", "A:got_s", "A:computed", "A:arrived", "B:top", "B:got_s", "B:computed", "B:arrived"]
print("MMA thread (t: issueS, issuedS, wait_p, got_p, issued dV/dK/dQ) relative cycles")
for t in range(0, 30):
    print(t, [int(tr[k, t]) - t0 if int(tr[k, t]) else None for k in (0, 11, 1, 2, 12)])
print("group A / B per n: top, got_s, computed, arrived")
for n in range(0, 14):
    print(n, [int(tr[k, n]) - t0 if int(tr[k, n]) else None for k in range(3, 11)])
print("group A detail per n: got_s, c0 ldwait, c0 math, c1 ldwait, c1 math, computed(before st wait), arrived")
for n in range(0, 14):
    print(n, [int(tr[k, n]) - t0 if int(tr[k, n]) else None for k in (4, 13, 14, 15, 16, 5, 6)])
last = max(int(tr[k].max()) for k in range(11)) - t0
print("total cycles", last)
