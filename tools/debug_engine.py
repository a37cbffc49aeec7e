"""Layer-by-layer comparison of the CUDA UNet engine with the oracle UNet (fp32 gold and fp16-autocast, both run on
the GPU for speed). usage: python tools/debug_engine.py [tiny|full] [B]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from instantrestore_b200 import _lib as L  # noqa: E402
from instantrestore_b200.unet_engine import UNetEngine, UNetSpec  # noqa: E402
from instantrestore_b200.weights import StateDictView  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.unet import UNetConfig  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = UNetConfig.tiny() if which == "tiny" else UNetConfig()
    spec = UNetSpec(block_out_channels=tuple(cfg.block_out_channels), attention_head_dim=tuple(cfg.attention_head_dim),
                    cross_attention_dim=cfg.cross_attention_dim)
    unet = synth.make_unet(cfg, seed=0, lora_rank=4).cuda()
    cap = synth.caption_embedding(cfg.cross_attention_dim)
    eng = UNetEngine(StateDictView(unet.state_dict()), spec, 249, cap, "cuda:0")
    S = cfg.sample_size
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 4, S, S, generator=g).cuda()
    acts = {}

    def hook(name):
        def f(mod, inp, out):
            o = out[0] if isinstance(out, tuple) else out
            acts[name] = o.detach().float()
        return f

    for name, m in unet.named_modules():
        parts = name.split(".")
        if name in ("conv_in", "mid_block") or (len(parts) == 4 and parts[2] in ("resnets", "attentions", "downsamplers", "upsamplers")):
            m.register_forward_hook(hook(name))
    t = torch.tensor([249], device="cuda")
    capd = cap.cuda().repeat(B, 1, 1)
    with torch.no_grad():
        gold = unet(x, t, encoder_hidden_states=capd)
        gold_acts = dict(acts)
        acts.clear()
        with torch.autocast("cuda", dtype=torch.float16):
            ac = unet(x, t, encoder_hidden_states=capd).float()
        ac_acts = dict(acts)
    eng.debug = {}
    xin = L.latent_in(x, None, 1.0, 0.0)
    out = eng.forward(xin, B, S, S)
    torch.cuda.synchronize()
    print(f"{'module':40s} {'shape':22s} {'ours-vs-fp32':>12s} {'autocast-vs-fp32':>16s}")
    for name, tt in eng.debug.items():
        ref = gold_acts[name]
        b, c, h, w = ref.shape
        ours = tt.float().view(b, h, w, c).permute(0, 3, 1, 2)
        print(f"{name:40s} {tuple(ref.shape)!s:22s} {rel(ours, ref):12.3e} {rel(ac_acts[name], ref):16.3e}")
    ours = out.float().view(B, S, S, 4).permute(0, 3, 1, 2)
    print(f"{'out':40s} {'':22s} {rel(ours, gold):12.3e} {rel(ac, gold):16.3e}")


if __name__ == "__main__":
    main()
