"""Developer aid: throughput of one finetune step (CAVMAEFT_BASE 'mm_grad', AudioSet 527 classes, BCE, FusedAdam)
on one GPU — BASELINE.json config 3 geometry (no masking: 512 + 196 tokens per sample)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsiam_b200 import CAVMAEFT_BASE, FusedAdam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = CAVMAEFT_BASE(label_dim=527).cuda()
model.direct_grads = True
opt = FusedAdam(model.parameters(), lr=1e-4, weight_decay=5e-7, betas=(0.95, 0.999), model=model)
a = torch.randn(B, 1024, 128, device="cuda"); v = torch.randn(B, 1, 3, 224, 224, device="cuda")
y = (torch.rand(B, 527, device="cuda") < 0.01).float()
bce = torch.nn.BCEWithLogitsLoss()
def step():
    out, out_a, out_v = model(a, v, "mm_grad")
    loss = bce(out, y) + bce(out_a, y) + bce(out_v, y)
    opt.zero_grad(); loss.backward(); opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"finetune mm_grad B={B}: {ms:.1f} ms/step = {B/ms*1e3:.0f} samples/s (loss {float(l):.4f})")
