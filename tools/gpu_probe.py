"""First-contact diagnostics on a B200: runs groups of checks in subprocesses (a trapped kernel poisons
its CUDA context) and prints one line per case.  Usage: python tools/gpu_probe.py [group ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_conv(x, w, bias, dilation, res, res_gain, out_slope):
    import torch
    import torch.nn.functional as F
    k = w.shape[2]
    xf = x.float().transpose(1, 2)
    y = F.conv1d(xf, w.bfloat16().float(), bias, dilation=dilation, padding=(k - 1) // 2 * dilation)
    if res is not None:
        r = res.float().transpose(1, 2)
        y = y + torch.where(r >= 0, r, r * res_gain)
    y = torch.where(y >= 0, y, y * out_slope)
    return y.transpose(1, 2)


def err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max()), float((a - b).abs().max() / (b.abs().max() + 1e-12))


def group_ops(impl, desc_mode, cases):
    import torch
    import vitsdec
    ops = __import__("importlib").import_module("personalized_text-to-speech_b200.ops")
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    for (B, L, ci, co, k, d, use_res) in cases:
        x = torch.randn(B, L, ci, device=dev).bfloat16()
        w = torch.randn(co, ci, k, device=dev) / (ci * k) ** 0.5
        b = torch.randn(co, device=dev) * 0.1
        res = torch.randn(B, L, co, device=dev).bfloat16() if use_res else None
        t0 = time.time()
        try:
            y = ops.conv1d_cl(x, w, b, dilation=d, res=res, res_gain=10.0, out_slope=0.1, impl=impl, desc_mode=desc_mode)
            torch.cuda.synchronize()
            e = err(y, ref_conv(x, w, b, d, res, 10.0, 0.1))
            print("conv impl=%d dm=%d B=%d L=%d ci=%d co=%d k=%d d=%d res=%d  maxabs %.3e rel %.3e  %s  (%.2fs)" % (
                impl, desc_mode, B, L, ci, co, k, d, use_res, e[0], e[1], "OK" if e[1] < 2e-2 else "MISMATCH",
                time.time() - t0), flush=True)
        except Exception as ex:
            print("conv impl=%d dm=%d %s EXC %s" % (impl, desc_mode, (B, L, ci, co, k, d, use_res), ex), flush=True)
            return


def group_convt(impl):
    import torch
    import torch.nn.functional as F
    import vitsdec
    ops = __import__("importlib").import_module("personalized_text-to-speech_b200.ops")
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    for (B, L, ci, co, k, s) in [(2, 50, 64, 32, 4, 2), (1, 37, 128, 64, 4, 2), (2, 33, 512, 256, 16, 8),
                                 (1, 130, 256, 128, 16, 8), (1, 9, 64, 32, 8, 4)]:
        x = torch.randn(B, L, ci, device=dev).bfloat16()
        w = torch.randn(ci, co, k, device=dev) / (ci * k / s) ** 0.5
        b = torch.randn(co, device=dev) * 0.1
        try:
            y = ops.conv_transpose1d_cl(x, w, b, stride=s, out_slope=0.1, impl=impl)
            torch.cuda.synchronize()
            r = F.conv_transpose1d(x.float().transpose(1, 2), w.bfloat16().float(), b, stride=s, padding=(k - s) // 2)
            r = torch.where(r >= 0, r, r * 0.1).transpose(1, 2)
            e = err(y, r)
            print("convT impl=%d B=%d L=%d ci=%d co=%d k=%d s=%d  maxabs %.3e rel %.3e %s" % (
                impl, B, L, ci, co, k, s, e[0], e[1], "OK" if e[1] < 2e-2 else "MISMATCH"), flush=True)
        except Exception as ex:
            print("convT impl=%d %s EXC %s" % (impl, (B, L, ci, co, k, s), ex), flush=True)
            return


def group_e2e(impl, hp_name, B, T):
    import numpy as np
    import torch
    import oracle
    import vitsdec
    from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
    hp = getattr(oracle, hp_name)
    sd = oracle.synth_state_dict(hp, 21, gain=2.0)
    args, kw = hp.ctor_args()
    G = vitsdec.Generator(*args, **kw)
    G.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    G = G.cuda().eval()
    G.set_option("impl", impl)
    G.set_option("debug_keep", 1)
    rs = np.random.RandomState(5)
    z = torch.from_numpy(rs.standard_normal((B, hp.initial_channel, T)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((B, hp.gin_channels, 1)).astype(np.float32)) if hp.gin_channels else None
    taps = {}
    ref = generator_forward_torch(hp, to_torch_state_dict(sd), z, g, taps=taps)
    with torch.no_grad():
        t0 = time.time()
        y = G(z.cuda(), None if g is None else g.cuda())
        torch.cuda.synchronize()
        dt = time.time() - t0
    y = y.cpu()
    num = float((ref ** 2).sum())
    den = float(((ref - y) ** 2).sum())
    snr = 10 * np.log10(num / max(den, 1e-30))
    print("e2e impl=%d %s B=%d T=%d  maxabs %.3e (ref max %.3e)  SNR %.1f dB  launches %d  %.2fs" % (
        impl, hp_name, B, T, float((ref - y).abs().max()), float(ref.abs().max()), snr, G.last_launch_count(), dt),
        flush=True)
    for name in taps:
        try:
            got = G.debug_read(name, B, T).cpu()
            r = taps[name]
            s = 10 * np.log10(float((r ** 2).sum()) / max(float(((r - got) ** 2).sum()), 1e-30))
            print("    %-10s SNR %.1f dB  max|ref| %.3e" % (name, s, float(r.abs().max())), flush=True)
        except Exception as ex:
            print("    %-10s EXC %s" % (name, ex), flush=True)


BASIC = [(1, 128, 64, 32, 1, 1, 0), (1, 128, 64, 64, 1, 1, 0), (1, 256, 64, 128, 1, 1, 0), (1, 128, 64, 256, 1, 1, 0),
         (1, 128, 128, 128, 1, 1, 0), (1, 128, 32, 32, 1, 1, 0)]
SHIFT = [(1, 128, 64, 64, 3, 8, 0), (1, 128, 64, 64, 3, 1, 0), (1, 200, 64, 64, 3, 3, 0), (2, 300, 128, 128, 7, 5, 1),
         (2, 1000, 256, 256, 11, 5, 1), (3, 777, 32, 32, 11, 3, 1), (2, 50, 192, 512, 7, 1, 0), (1, 5, 64, 64, 7, 1, 1),
         (1, 4000, 64, 64, 11, 5, 1)]

GROUPS = {
    "simt_ops": lambda: group_ops(1, 0, BASIC + SHIFT),
    "simt_convt": lambda: group_convt(1),
    "simt_e2e_tiny": lambda: group_e2e(1, "TINY", 2, 9),
    "simt_e2e_full": lambda: group_e2e(1, "FINETUNE_SPEAKER", 2, 32),
    "tc_basic": lambda: group_ops(0, 0, BASIC),
    "tc_shift_dm0": lambda: group_ops(0, 0, SHIFT),
    "tc_shift_dm1": lambda: group_ops(0, 1, SHIFT),
    "tc_convt": lambda: group_convt(0),
    "tc_e2e_tiny": lambda: group_e2e(0, "TINY", 2, 9),
    "tc_e2e_full": lambda: group_e2e(0, "FINETUNE_SPEAKER", 2, 32),
    "tc_e2e_full_big": lambda: group_e2e(0, "FINETUNE_SPEAKER", 3, 173),
}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        GROUPS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(GROUPS)
    for n in names:
        print("=== %s" % n, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], timeout=240,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-6000:], flush=True)
            print("--- exit %d" % r.returncode, flush=True)
        except subprocess.TimeoutExpired as e:
            print((e.stdout or b"").decode(errors="replace")[-3000:] if isinstance(e.stdout, bytes) else (e.stdout or ""))
            print("--- TIMEOUT", flush=True)
