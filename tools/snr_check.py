"""SNR of the decoder vs the fp32 restatement for the schedule variants (fold / pairf options)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, vitsdec
from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
hp = oracle.FINETUNE_SPEAKER
def snr(ref, y): return 10 * np.log10(float((ref ** 2).sum()) / float(((ref - y) ** 2).sum()))
for seed, B, T in ((21, 2, 20), (36, 2, 173), (44, 1, 400)):
    sd = oracle.synth_state_dict(hp, seed, gain=2.0)
    args, kw = hp.ctor_args()
    G = vitsdec.Generator(*args, **kw)
    G.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    G = G.to("cuda:0").eval()
    rs = np.random.RandomState(5)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32))
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, g)
    out = []
    for opts in ({"fold": 0, "pairf": 0}, {"fold": 1, "pairf": 0}, {"fold": 1, "pairf": 1}, {"fold": 1, "pairf": 2}, {"fold": 1, "pairf": 1, "fuse_pairs": 0}):
        for k_, v in {"fuse_pairs": 1, **opts}.items():
            G.set_option(k_, v)
        with torch.no_grad():
            y = G(z.cuda(), g.cuda()).cpu()
        out.append("%s %.2f dB" % (opts, snr(ref, y)))
    print((seed, B, T), " | ".join(out))
    for k_, v in {"fuse_pairs": 1, "fold": 1, "pairf": 1}.items():
        G.set_option(k_, v)
    out = []
    for f16 in (0, 1):   # 16-bit storage format of weights and activations
        G.set_option("fp16", f16)
        with torch.no_grad():
            y = G(z.cuda(), g.cuda()).cpu()
        out.append("fp16=%d %.2f dB, max-abs %.5f of peak" % (f16, snr(ref, y), float((ref - y).abs().max() / ref.abs().max())))
    print((seed, B, T), " | ".join(out))
