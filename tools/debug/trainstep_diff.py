"""debug: TrainStep vs two stock loops (noise floor) in bf16"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tfswa_unet_b200 as T
from tfswa_unet_b200.train_step import TrainStep, masked_magnitude_l1
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
T.set_precision(prec)
torch.manual_seed(0)
mk = lambda: T.TFSWAUNet(4, 4, [1, 1, 1, 1], [32, 64, 128, 256], 8, 4, 8).train().cuda()
a, b, c = mk(), mk(), mk()
b.load_state_dict(a.state_dict()); c.load_state_dict(a.state_dict())
ts = TrainStep(a, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
ob = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=1e-2)
oc = torch.optim.AdamW(c.parameters(), lr=1e-3, weight_decay=1e-2)
g = torch.Generator(device="cuda").manual_seed(2)
def top(m1, m2, what):
    rows = []
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    for k in sd1:
        if what(k):
            u, v = sd1[k].float(), sd2[k].float()
            rows.append((float(((u - v).abs() / (v.abs() + 1e-2)).max()), float((u - v).abs().max()), float(v.abs().max()), k))
    rows.sort(reverse=True)
    return rows[:4]
for step in range(2):
    x = torch.randn(2, 4, 40, 24, device="cuda", generator=g)
    mix = torch.rand(2, 40, 24, device="cuda", generator=g)
    tg = [torch.rand(2, 40, 24, device="cuda", generator=g) for _ in range(2)]
    la, na = ts(x, mix, tg)
    for m, o in ((b, ob), (c, oc)):
        o.zero_grad()
        l = masked_magnitude_l1(m(x), mix, tg); l.backward()
        n = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0); o.step()
        print(f"step {step} stock loss {float(l):.7f} norm {float(n):.6f}")
    print(f"step {step} ts    loss {float(la):.7f} norm {float(na):.6f}")
    for name, (m1, m2) in {"a-b": (a, b), "c-b": (c, b)}.items():
        print(name, "running:", top(m1, m2, lambda k: "running" in k))
        print(name, "params :", top(m1, m2, lambda k: "running" not in k and "num_batches" not in k and "attn_mask" not in k))
for m in (a, b, c): m.eval()
with torch.no_grad():
    ya, yb, yc = a(x), b(x), c(x)
print("eval a-b", float((ya - yb).abs().max()), "c-b", float((yc - yb).abs().max()))
with torch.no_grad():
    ya2 = a(x)
print("eval a again", float((ya2 - ya).abs().max()))
