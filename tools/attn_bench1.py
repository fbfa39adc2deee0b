"""Developer aid: time attention bwd on the decoder shape only."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import ops
n_seq, S, H, hd = 256, 708, 16, 32
D = H * hd
qkv = torch.randn(n_seq * S, 3 * D, device="cuda").bfloat16()
out = torch.empty(n_seq * S, D, device="cuda", dtype=torch.bfloat16)
dout = torch.randn_like(out); lse = torch.zeros(n_seq, H, S, device="cuda"); delta = torch.empty_like(lse); dqkv = torch.empty_like(qkv)
ops.attention_fwd(qkv, out, lse, n_seq, S, H, hd)
def run(): ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, n_seq, S, H, hd)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"AVS_TC_EXP={os.environ.get('AVS_TC_EXP','0')}: bwd {e0.elapsed_time(e1)/10:.3f} ms")
