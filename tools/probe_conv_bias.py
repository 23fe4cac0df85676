"""How does eager round `F.conv2d(x, w, bias)` on CUDA under bf16?  (a) ONE rounding, bf16(acc + bias), like F.linear;
(b) TWO: the convolution output is rounded to bf16 and the bias is added by a second bf16 op (ATen's cuDNN path).
Prints the fraction of outputs bit-identical to each model, for the ViT patch-embed stem and an SD-v1.5 3x3 convolution."""
import torch
import torch.nn.functional as F

dev = "cuda"
g = torch.Generator().manual_seed(0)
for name, (B, Cin, Cout, S, k, stride, pad) in {"vit stem 16x16/16": (8, 3, 768, 224, 16, 16, 0), "sd conv 3x3": (2, 320, 320, 64, 3, 1, 1)}.items():
    x = torch.randn(B, Cin, S, S, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (1.0 / (Cin * k * k) ** 0.5)).to(dev).to(torch.bfloat16)
    b = (0.5 * torch.randn(Cout, generator=g)).to(dev).to(torch.bfloat16)
    eager = F.conv2d(x, w, b, stride=stride, padding=pad)
    nobias = F.conv2d(x, w, None, stride=stride, padding=pad)
    two = nobias + b.view(1, -1, 1, 1)
    acc = F.conv2d(x.float(), w.float(), None, stride=stride, padding=pad)
    one = (acc + b.float().view(1, -1, 1, 1)).to(torch.bfloat16)
    two_f = (acc.to(torch.bfloat16).float() + b.float().view(1, -1, 1, 1)).to(torch.bfloat16)
    eq = lambda a, c: (a == c).float().mean().item()  # noqa: E731
    print(f"{name}: eager == [conv_bf16 + bias in bf16] {eq(eager, two):.6f} | eager == bf16(bf16(fp32 conv) + bias) {eq(eager, two_f):.6f} | "
          f"eager == bf16(fp32 conv + bias) {eq(eager, one):.6f}")
