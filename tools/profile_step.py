"""Runs a few device-resident ViT-B/16 training steps (B=256) for ncu: `ncu ... python tools/profile_step.py [steps]`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cflearn_b200  # noqa: F401,E402
from cflearn_b200 import registry  # noqa: E402
from cflearn_b200.optim import ArenaAdam  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=1000, img_size=224, latent_dim=768, encoder="vit",
                                                encoder_config=dict(patch_size=16, num_layers=12))).to(dev)
opt = ArenaAdam(m)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (B, 1), device=dev)
for _ in range(steps):
    opt.zero_grad()
    loss = m.train_step(x, y)
    opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
