import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swin_b200
from swin_b200 import ops, _lib as L
from oracle import swin_oracle as so
torch.manual_seed(0)
for (H, W, C) in [(56, 56, 96), (200, 334, 96), (100, 167, 192)]:
    x = torch.randn(1, H * W, C)
    g = 1 + 0.2 * torch.randn(C); b = 0.3 * torch.randn(C)
    ref = so.layer_norm(x.double(), g.double(), b.double())
    y, m, r = ops.ln_fwd(0, x.cuda(), g.cuda(), b.cuda(), 1, H * W, 1, C, 1, 0, 1e-5, L.F32)
    print(H, W, C, "ln mode0 f32:", so.rel_l2(y.view(1, H * W, C), ref))
    out, m, r = ops.ln_nchw_fwd(x.cuda(), g.cuda(), b.cuda(), H, W, 1e-5)
    want = ref.reshape(1, H, W, C).permute(0, 3, 1, 2)
    e = (out.cpu().double() - want)
    print(H, W, C, "ln_nchw:", so.rel_l2(out, want), "per-channel err max", e.flatten(2).norm(dim=2).max().item())
# full patch embed module vs oracle at 800x1333
from oracle.make_golden import rnd
shapes = {k: v for k, v in so.param_shapes(96, [2], [3], out_indices=(0,)).items() if k.startswith("patch_embed")}
params = so.seeded_params(shapes, seed=7)
for mode in ("fp32", "bf16"):
    pe = swin_b200.PatchEmbed(4, 3, 96, torch.nn.LayerNorm, compute_dtype=mode)
    sd = pe.state_dict()
    for k in sd: sd[k] = params["patch_embed." + k]
    pe.load_state_dict(sd); pe = pe.cuda()
    for hw in ((224, 224), (800, 1333)):
        img = torch.from_numpy(rnd(1, (1, 3) + hw))
        tok, Wh, Ww = pe.tokens(img.cuda())
        ref, _, _ = so.patch_embed(img.double(), {k: v.double() for k, v in params.items()}, 4, True)
        print(mode, hw, "patch_embed tokens err", so.rel_l2(tok, ref), Wh, Ww)
