import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swin_b200
from oracle import swin_oracle as so
from oracle.make_golden import rnd
SWIN_T = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7)
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
HW = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (800, 1333)
params = so.seeded_params(so.param_shapes(**SWIN_T), seed=7)
net = swin_b200.SwinTransformer(drop_path_rate=0.0, compute_dtype=mode, **SWIN_T)
sd = net.state_dict()
for k in sd:
    if not k.endswith("relative_position_index"): sd[k] = params[k]
net.load_state_dict(sd); net = net.cuda().train()
img = torch.from_numpy(rnd(1, (1, 3) + HW))
torch.set_num_threads(os.cpu_count())
with torch.no_grad():
    # oracle intermediates
    x, H, W = so.patch_embed(img, params, 4, True)
    ours, Wh, Ww = net.patch_embed.tokens(img.cuda())
    print("tokens", so.rel_l2(ours, x), H, W, Wh, Ww)
    xo = ours
    for s in range(2):
        layer = net.layers[s]
        mask = layer.attn_mask(H, W, xo.device)
        for b, blk in enumerate(layer.blocks):
            shift = 0 if b % 2 == 0 else 3
            x = so.swin_block(x, H, W, params, f"layers.{s}.blocks.{b}.", SWIN_T["num_heads"][s], 7, shift)
            blk.H, blk.W = H, W
            xo = blk(xo, mask)
            print(f"stage {s} block {b}", so.rel_l2(xo, x))
        n = getattr(net, f"norm{s}")
        from swin_b200.functional import OutNormFn
        o = OutNormFn.apply(xo, n.weight, n.bias, H, W, 1e-5)
        r = so.layer_norm(x, params[f"norm{s}.weight"], params[f"norm{s}.bias"]).reshape(1, H, W, -1).permute(0, 3, 1, 2)
        print(f"out{s}", so.rel_l2(o, r))
        x = so.patch_merging(x, H, W, params, f"layers.{s}.downsample.")
        xo = layer.downsample(xo, H, W)
        H, W = (H + 1) // 2, (W + 1) // 2
        print(f"merge {s}", so.rel_l2(xo, x))
    outs = net(img.cuda())
    outs_r = so.backbone_forward(img, params, **SWIN_T)
    print("full:", [so.rel_l2(a, b) for a, b in zip(outs, outs_r)])
    outs = net(img.cuda())
    print("full again:", [so.rel_l2(a, b) for a, b in zip(outs, outs_r)])
