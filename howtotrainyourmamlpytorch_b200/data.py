"""GPU-resident episode sampler.

The reference builds every task on the host: ``FewShotLearningDatasetParallel.get_set`` (``data.py:478-524``) draws the
classes / samples / rotations from ``np.random.RandomState(seed)`` and then loads, rotates (PIL / NumPy) and stacks the
images in DataLoader worker processes.  At thousands of tasks per second that loader is the bottleneck, so here the
dataset lives in HBM and only the seeded index arithmetic stays on the host:

  * ``episode_indices(seed, augment)`` replays the reference's exact RNG call sequence (``choice`` of classes,
    ``shuffle``, ``randint`` rotations, per-class ``choice`` of samples) -- same seeds => same episodes;
  * ``sample_batch(seeds, augment)`` ships the few hundred indices with one small H2D copy and ONE kernel
    (``maml_b200_episode_gather``) gathers, rotates (``np.rot90``), normalises and writes the NCHW support / target
    tensors and class-major labels on the device -- ready for ``MAMLFewShotClassifier.run_train_iter``.

Dataset form = what the reference keeps in RAM when ``load_into_memory`` is set: per class an array ``[n, H, W, C]`` of
float32 (Omniglot: binary floats, no rescaling; ImageNet: ``x / 255``).  Reading image files is out of scope.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import _native

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class GpuEpisodeSampler(object):
    def __init__(self, args, class_images, device, dataset_name=None):
        """``class_images``: ordered mapping class key -> ndarray / tensor ``[n_c, H, W, C]`` float32 (the order is the
        reference's ``dataset_size_dict[set].keys()`` order)."""
        self.args = args
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _native.NativeLibraryError("GpuEpisodeSampler needs a CUDA device: no CPU fallback")
        self.dataset_name = dataset_name if dataset_name is not None else args.dataset_name
        self.N = int(args.num_classes_per_set)
        self.K = int(args.num_samples_per_class)
        self.T = int(args.num_target_samples)
        self.keys = list(class_images.keys())
        self.sizes = [int(len(class_images[k])) for k in self.keys]
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        first = np.asarray(class_images[self.keys[0]])
        self.H, self.W, self.C = int(first.shape[1]), int(first.shape[2]), int(first.shape[3])
        flat = np.concatenate([np.asarray(class_images[k], dtype=np.float32) for k in self.keys], axis=0)
        self.dataset = torch.from_numpy(np.ascontiguousarray(flat)).to(self.device)
        self._normalise = "imagenet" in self.dataset_name
        self._rotate = "omniglot" in self.dataset_name
        self._pin = {}

    # ------------------------------------------------------------------ host: the reference's RNG sequence
    def episode_indices(self, seed, augment):
        """(image_index [N, K+T] int64 rows of the flat dataset, rot_k [N] int32, class keys in episode-label order)
        for one task -- the draw sequence of reference ``get_set`` (data.py:484-503)."""
        rng = np.random.RandomState(seed)
        sel = rng.choice(len(self.keys), size=self.N, replace=False)        # == rng.choice(list(keys), ...): same permutation
        rng.shuffle(sel)
        k_list = rng.randint(0, 4, size=self.N)
        idx = np.empty((self.N, self.K + self.T), dtype=np.int64)
        for n, cls in enumerate(sel):
            chosen = rng.choice(self.sizes[cls], size=self.K + self.T, replace=False)
            idx[n] = self.offsets[cls] + chosen
        rot = (k_list if (augment and self._rotate) else np.zeros(self.N)).astype(np.int32)
        return idx, rot, [self.keys[c] for c in sel]

    # ------------------------------------------------------------------ device: gather + transform
    def sample_batch(self, seeds, augment=False):
        """Episode batch for ``seeds`` (one task each) as device tensors ``(x_support [B,N,K,C,H,W], x_target
        [B,N,T,C,H,W], y_support [B,N,K] int64, y_target [B,N,T] int64)`` -- the reference's DataLoader batch layout."""
        B = len(seeds)
        if self._rotate and augment and self.H != self.W:
            raise ValueError("rot90 augmentation needs square images")
        key = B
        st = self._pin.get(key)
        if st is None:
            nidx = B * self.N * (self.K + self.T)
            st = {"pin_idx": torch.empty(nidx, dtype=torch.int64).pin_memory(),
                  "pin_rot": torch.empty(B * self.N, dtype=torch.int32).pin_memory(),
                  "dev_idx": torch.empty(nidx, dtype=torch.int64, device=self.device),
                  "dev_rot": torch.empty(B * self.N, dtype=torch.int32, device=self.device)}
            self._pin[key] = st
        idx_np = st["pin_idx"].numpy().reshape(B, self.N, self.K + self.T)
        rot_np = st["pin_rot"].numpy().reshape(B, self.N)
        for b, seed in enumerate(seeds):
            idx_np[b], rot_np[b], _ = self.episode_indices(int(seed), augment)
        with torch.cuda.device(self.device):
            st["dev_idx"].copy_(st["pin_idx"], non_blocking=True)
            st["dev_rot"].copy_(st["pin_rot"], non_blocking=True)
            xs = torch.empty(B, self.N, self.K, self.C, self.H, self.W, dtype=torch.float32, device=self.device)
            xt = torch.empty(B, self.N, self.T, self.C, self.H, self.W, dtype=torch.float32, device=self.device)
            ys = torch.empty(B, self.N, self.K, dtype=torch.int64, device=self.device)
            yt = torch.empty(B, self.N, self.T, dtype=torch.int64, device=self.device)
            _native.episode_gather(self.dataset, st["dev_idx"], st["dev_rot"], B, self.N, self.K, self.T, self.C, self.H, self.W,
                                   IMAGENET_MEAN[:self.C] if self._normalise else None,
                                   IMAGENET_STD[:self.C] if self._normalise else None, xs, xt, ys, yt)
        return xs, xt, ys, yt


def synthetic_class_images(num_classes, samples_per_class, height, width, channels, seed=0, binary=False):
    """A seeded stand-in dataset in the reference's in-memory form (tests / demos)."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for c in range(num_classes):
        x = rng.rand(samples_per_class, height, width, channels).astype(np.float32)
        out["class_%03d" % c] = (x < 0.93).astype(np.float32) if binary else x
    return out
