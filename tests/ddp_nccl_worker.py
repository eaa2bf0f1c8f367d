"""Worker of test_two_gpu_nccl_sharded_gradients_equal_full_batch (launched by torchrun, one rank per GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import swin_b200
    from swin_b200.ddp import BucketedGradAllReduce
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=7, out_indices=(0, 1), drop_path_rate=0.0)
    worst = 0.0
    for mode, tol in (("fp32", 1e-5), ("bf16", 2e-2)):
        torch.manual_seed(0)
        net = swin_b200.SwinTransformer(compute_dtype=mode, **cfg)
        net.init_weights()
        if rank == 1:                                  # rank 1 starts from other weights: the constructor broadcast must fix it
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(0.5)
        net = net.to(dev).train()
        if rank == 1:
            with torch.no_grad():
                net(torch.randn(1, 3, 56, 56, device=dev))   # populates the bf16 operand cache with the WRONG weights first
        ddp = BucketedGradAllReduce(net, bucket_mb=0.05, tail_kb=4.0)
        img = torch.from_numpy(np.random.default_rng(0).standard_normal((4, 3, 56, 84)).astype(np.float32)).to(dev)
        cots = None
        for it in range(3):                            # step 0 measures the arrival order and rebuilds the buckets
            ddp.zero_grad()
            outs = net(img[rank * 2:(rank + 1) * 2])
            if cots is None:
                g = torch.Generator(device=dev).manual_seed(3)
                full = [torch.randn((4,) + tuple(o.shape[1:]), device=dev, generator=g) for o in outs]
                cots = [c[rank * 2:(rank + 1) * 2] * world for c in full]      # AVG over ranks of world * shard-sum == full sum
            torch.autograd.backward(outs, cots)
            ddp.finish()
            torch.cuda.synchronize()
        assert ddp._ordered and len(ddp.buckets) > 2
        grads = {k: v.grad.clone() for k, v in net.named_parameters()}
        # the single-GPU full batch on the same (rank 0) weights
        ddp.zero_grad()
        outs = net(img)
        torch.autograd.backward(outs, full)
        torch.cuda.synchronize()
        for k, v in net.named_parameters():
            den = v.grad.double().norm().item() or 1.0
            e = (grads[k].double() - v.grad.double()).norm().item() / den
            worst = max(worst, e)
            assert e < tol, (mode, k, e)
    dist.barrier()
    if rank == 0:
        print(f"DDP_NCCL_OK worst rel-L2 {worst:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
