"""torchrun --nproc-per-node 2 tools/dp_check.py : data-parallel gradient parity on real GPUs.

Each rank runs the B200 training step on its shard (DistributedSampler rule: rank r takes samples r, r+W, ...);
the bucketed NCCL all-reduce(AVG) must reproduce the gradients of the global-batch mean loss, which rank 0 also
computes alone on the concatenated batch (SURVEY.md section 8e / finding 3: the reference defines DP parity mathematically)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vit_oracle as vo  # noqa: E402

import cflearn_b200  # noqa: F401,E402
from cflearn_b200 import dp, registry  # noqa: E402


def build(cfg, sd, dev):
    m = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=cfg["num_classes"], img_size=cfg["img_size"],
                                                    latent_dim=cfg["latent_dim"], encoder="vit",
                                                    encoder_config=dict(patch_size=cfg["patch_size"], num_layers=cfg["num_layers"])))
    m.load_state_dict(sd, strict=True)
    return m.to(dev)


def main():
    rank, world, local = dp.ddp_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = vo.vit_config("vit_small")
    sd = vo.init_state_dict(cfg, seed=0)
    per_rank = 8
    x, y = vo.synthetic_batch(cfg, per_rank * world, seed=3)
    idx = list(dp.shard_indices(per_rank * world, rank, world))
    m = build(cfg, sd, dev)
    dp.broadcast_parameters(m)
    dp.attach_reducer(m)
    loss = m.train_step(x[idx].to(dev), y[idx].to(dev))
    torch.cuda.synchronize()
    g_dp = m.arena.grad.clone()
    # every rank must hold the same averaged gradients
    ref = g_dp.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, g_dp), "ranks disagree after the all-reduce"
    if rank == 0:
        single = build(cfg, sd, dev)
        single.train_step(x.to(dev), y.to(dev))
        torch.cuda.synchronize()
        g_one = single.arena.grad
        rel = ((g_dp - g_one).norm() / g_one.norm()).item()
        # same bf16 noise-floor argument as tests/test_model_gpu.py: shard-wise bf16 rounding of the weight grads
        print(f"dp_check: world {world}, loss(rank0 shard) {loss.item():.5f}, rel L2 (all-reduced vs single-process global batch) = {rel:.3e}")
        assert rel < 2e-2, rel
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
