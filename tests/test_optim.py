"""Row f3 (fused AdamW).  CPU: the paramwise weight-decay rule of the reference configs.  GPU: the fused multi-tensor
kernel against torch.optim.AdamW on the same seeded gradients (fp32, tolerance 2e-6 relative), including the bf16
shadow refresh, ragged sizes and more tensors than one launch takes."""
import pytest
import torch


class _Blk(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.norm1 = torch.nn.LayerNorm(8)
        self.fc = torch.nn.Linear(8, 8)
        self.relative_position_bias_table = torch.nn.Parameter(torch.zeros(9, 2))
        self.absolute_pos_embed = torch.nn.Parameter(torch.zeros(1, 8, 2, 2))
        self.frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)


def test_paramwise_weight_decay_rule():
    """configs/swin/mask_rcnn_swin_tiny_patch4_window7_mstrain_480-800_adamw_1x_coco.py:64-67: decay_mult 0 for names containing
    'absolute_pos_embed', 'relative_position_bias_table', 'norm'; base decay elsewhere; frozen parameters are skipped."""
    from swin_b200.optim import paramwise_weight_decay
    m = _Blk()
    got = {id(p): wd for p, wd in paramwise_weight_decay(m.named_parameters(), 0.05)}
    names = dict(m.named_parameters())
    assert got[id(names["norm1.weight"])] == 0.0 and got[id(names["norm1.bias"])] == 0.0
    assert got[id(names["relative_position_bias_table"])] == 0.0 and got[id(names["absolute_pos_embed"])] == 0.0
    assert got[id(names["fc.weight"])] == 0.05 and got[id(names["fc.bias"])] == 0.05
    assert id(names["frozen"]) not in got
    half = {id(p): wd for p, wd in paramwise_weight_decay(m.named_parameters(), 0.1, {"fc": 0.5})}
    assert half[id(names["fc.weight"])] == pytest.approx(0.05) and half[id(names["norm1.weight"])] == 0.1


def test_swin_t_decay_groups():
    """On the real backbone the rule leaves decay on exactly the Linear / conv weights and biases."""
    import swin_b200
    from swin_b200.optim import paramwise_weight_decay
    net = swin_b200.SwinTransformer()
    pairs = paramwise_weight_decay(net.named_parameters(), 0.05)
    name_of = {id(p): n for n, p in net.named_parameters()}
    for p, wd in pairs:
        n = name_of[id(p)]
        no_decay = ("norm" in n) or ("relative_position_bias_table" in n)
        assert wd == (0.0 if no_decay else 0.05), n
    assert len(pairs) == len(list(net.parameters()))


@pytest.mark.gpu
def test_fused_adamw_matches_torch():
    from swin_b200 import ops
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 96, 169 * 3, 4096, 4097, 288 * 96] + [int(x) for x in torch.randint(1, 3000, (70,), generator=g)]
    params = [torch.randn(n, generator=g).to(dev) for n in sizes]
    # a view at an odd element offset: its pointers are not 16-byte aligned, so the kernel's scalar path updates it
    sizes.append(1001)
    params.append(torch.randn(1004, generator=g).to(dev)[1:1002])
    wds = [0.05 if i % 3 else 0.0 for i in range(len(sizes))]
    ref_params = [torch.nn.Parameter(p.clone()) for p in params]
    ref = torch.optim.AdamW([{"params": [rp], "weight_decay": wd} for rp, wd in zip(ref_params, wds)], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    ms = [torch.zeros_like(p) for p in params]
    vs = [torch.zeros_like(p) for p in params]
    shadows = [p.bfloat16() if i % 2 == 0 else None for i, p in enumerate(params)]
    for step in range(1, 4):
        grads = [torch.randn(n, generator=g).to(dev) for n in sizes]
        for rp, gr in zip(ref_params, grads):
            rp.grad = gr.clone()
        ref.step()
        ops.adamw_step(params, grads, ms, vs, shadows, wds, 1e-3, 0.9, 0.999, 1e-8, step)
    for p, rp, s in zip(params, ref_params, shadows):
        assert torch.allclose(p, rp.detach(), rtol=2e-6, atol=1e-7)
        if s is not None:
            assert torch.equal(s, p.bfloat16())
    st = ref.state[ref_params[3]]
    assert torch.allclose(ms[3], st["exp_avg"], rtol=1e-5, atol=1e-7) and torch.allclose(vs[3], st["exp_avg_sq"], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_fused_adamw_trains_backbone_and_keeps_shadows_fresh():
    """One optimizer step on a small backbone: the bf16 operand copies the GEMMs read are refreshed by the fused kernel
    (outputs change, and equal a model rebuilt from the updated fp32 weights)."""
    import swin_b200
    from swin_b200.optim import FusedAdamW
    torch.manual_seed(0)
    kw = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], out_indices=(0, 1), drop_path_rate=0.0)
    net = swin_b200.SwinTransformer(**kw).cuda().train()
    x = torch.randn(2, 3, 56, 84, device="cuda")
    opt = FusedAdamW(net, lr=1e-2, weight_decay=0.05)
    out0 = [o.clone() for o in net(x)]
    sum(o.square().mean() for o in net(x)).backward()
    opt.step()
    out1 = net(x)
    assert not torch.allclose(out0[1], out1[1])
    twin = swin_b200.SwinTransformer(**kw).cuda().train()
    twin.load_state_dict(net.state_dict())
    for a, b in zip(out1, twin(x)):
        assert torch.equal(a, b)
