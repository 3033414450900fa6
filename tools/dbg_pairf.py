import importlib, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitsdec
ops = importlib.import_module("personalized_text-to-speech_b200.ops")
dev = torch.device("cuda:0")
def ref_pair(x, w1, b1, w2, b2, d, slope=0.1):
    k = w1.shape[2]
    a = x.float().transpose(1, 2)
    h = F.conv1d(a, w1.bfloat16().float(), b1, dilation=d, padding=(k - 1) // 2 * d)
    h = torch.where(h >= 0, h, h * slope).bfloat16().float()
    y = F.conv1d(h, w2.bfloat16().float(), b2, padding=(k - 1) // 2)
    y = y + torch.where(a >= 0, a, a / slope)
    return torch.where(y >= 0, y, y * slope).transpose(1, 2)
for (B, L, C, k, d) in [(2, 4000, 32, 3, 3), (3, 3108, 32, 7, 3), (2, 5000, 32, 11, 3), (1, 20, 32, 11, 3), (2, 3000, 64, 3, 3), (1, 4008, 32, 3, 3), (1, 4004, 32, 3, 3)]:
    torch.manual_seed(L + k + d)
    x = torch.randn(B, L, C, device=dev).bfloat16()
    w1 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    w2 = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    b1 = torch.randn(C, device=dev) * 0.1
    b2 = torch.randn(C, device=dev) * 0.1
    y = ops.resblock_pair_cl(x, w1, b1, w2, b2, dilation=d, slope=0.1, folded=True).float()
    r = ref_pair(x, w1, b1, w2, b2, d)
    err = (y - r).abs().amax(dim=2)  # [B, L]
    bad = (err > 0.05) | ~torch.isfinite(err)
    idx = bad.nonzero()
    print((B, L, C, k, d), "bad rows:", idx.shape[0], "first", idx[:3].tolist(), "last", idx[-3:].tolist())
