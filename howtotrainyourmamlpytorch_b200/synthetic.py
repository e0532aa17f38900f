"""Seeded synthetic episodes with the reference's batch layout.

Omniglot-shaped (C=1): Bernoulli(0.93) pixels in {0,1} (``kind='bernoulli'``), Mini-ImageNet-shaped:
N(0,1) (``kind='normal'``).  Labels ``y[b, c, :] = c`` as produced by reference ``data.py:491-514``.
Returned tensors are CPU float32 and contiguous, like what the reference's DataLoader yields.
"""
import torch


def synthetic_batch(args, iteration=0, batch_size=None, kind=None):
    B = int(batch_size if batch_size is not None else args.batch_size)
    N, K, T = int(args.num_classes_per_set), int(args.num_samples_per_class), int(args.num_target_samples)
    C, H, W = int(args.image_channels), int(args.image_height), int(args.image_width)
    gen = torch.Generator().manual_seed(1234 + int(iteration))
    if kind is None:
        kind = "bernoulli" if C == 1 else "normal"
    if kind == "bernoulli":
        xs = (torch.rand(B, N, K, C, H, W, generator=gen) < 0.93).float()
        xt = (torch.rand(B, N, T, C, H, W, generator=gen) < 0.93).float()
    else:
        xs = torch.randn(B, N, K, C, H, W, generator=gen)
        xt = torch.randn(B, N, T, C, H, W, generator=gen)
    ys = torch.arange(N).view(1, N, 1).expand(B, N, K).contiguous().float()
    yt = torch.arange(N).view(1, N, 1).expand(B, N, T).contiguous().float()
    return xs, xt, ys, yt
