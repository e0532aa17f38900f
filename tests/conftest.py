import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
TINY_CASES = ["tiny_pp", "tiny_pp_late", "tiny_pp_first", "tiny_maml", "tiny_odd", "tiny_bern"]
BIG_CASES = ["omniglot_mamlpp_5w1s", "omniglot_maml_5w1s", "mini_imagenet_mamlpp_5w1s", "omniglot_mamlpp_20w5s",
             "omniglot_mamlpp_5w1s_bernoulli", "mini_imagenet_mamlpp_5w5s"]
# Bernoulli(0.93) binary images (the distribution bench.py runs): exact max-pool ties, resolved first-max-wins
BERNOULLI_CASES = ["tiny_bern", "omniglot_mamlpp_5w1s_bernoulli"]
ALL_CASES = TINY_CASES + BIG_CASES


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device")


class Golden(object):
    """One tests/golden/<case>.npz produced by oracle/gen_golden.py from the unmodified reference."""

    def __init__(self, case):
        from howtotrainyourmamlpytorch_b200.utils.parser_utils import args_from_json
        self.case = case
        self.blob = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
        self.argdict = json.loads(str(self.blob["args_json"]))
        self.args = args_from_json(None, **self.argdict)
        self.iters = json.loads(str(self.blob["iters_json"]))
        self.kind = str(self.blob["kind"])

    def state(self, dtype=torch.float32):
        return {k[len("state/"):]: torch.from_numpy(self.blob[k]).to(dtype) for k in self.blob.files
                if k.startswith("state/")}

    def batch(self, it=0):
        from oracle import maml_oracle as O
        key = "it%d/xs" % it
        if key in self.blob.files:
            return tuple(torch.from_numpy(self.blob["it%d/%s" % (it, n)]) for n in ("xs", "xt", "ys", "yt"))
        return O.synthetic_batch(self.args, iteration=self.iters[it][1], kind=self.kind)

    def grads(self, it=0, suffix=""):
        pre = "it%d/grad%s/" % (it, suffix)
        return {k[len(pre):]: torch.from_numpy(self.blob[k]) for k in self.blob.files if k.startswith(pre)}

    def post(self, it=0):
        pre = "it%d/post/" % it
        return {k[len(pre):]: torch.from_numpy(self.blob[k]) for k in self.blob.files if k.startswith(pre)}

    def scalar(self, name, it=0):
        return float(self.blob["it%d/%s" % (it, name)])

    def val(self, name):
        """Reference ``run_validation_iter`` outputs from the initial state on batch 0 (``val/<name>``)."""
        return self.blob["val/" + name]

    def val_post(self):
        return {k[len("val/post/"):]: torch.from_numpy(self.blob[k]) for k in self.blob.files if k.startswith("val/post/")}

    def array(self, name, it=0):
        return self.blob["it%d/%s" % (it, name)]


_cache = {}


def load_golden(case):
    if case not in _cache:
        _cache[case] = Golden(case)
    return _cache[case]


def grad_tolerance(name, ref32, ref64, big=False):
    """Tolerance policy: allowed |X - X_ref64|_inf for one tensor.

    3x the reference's own fp32-vs-fp64 distance (5x for full-size cases, see below), floored at `rel` of the
    tensor's max-norm:
      * tiny cases: rel = 2e-5 (pure fp32 rounding; ~10 inner-loop-amplified ulps);
      * full-size cases: rel = 2e-2.  With 10^5..10^7 activations per pass some pre-activation sits
        within one fp32 ulp of 0 (leaky-ReLU branch) or of its pooling neighbour (arg-max), and ANY
        change of summation order flips that discrete choice.  One flip moves the meta-gradient by
        ~1e-4 of its max-norm (measured: the autograd-free fp32 restatement vs the fp64 reference on
        omniglot_mamlpp_5w1s, 3 flips -> 4.6e-4; the GPU path, other flips -> up to 3.6e-3 on one LSLR
        gradient; DESIGN.md "noise floor").  The TIGHT full-size check is
        tests/test_gpu_parity.py::test_decision_forced_parity, which pins the discrete decisions
        and then demands 1e-4; stage-level tests stay at 1e-5.
        Where the reference's own fp32 run already sits 10-30 % of max-norm away from its fp64 run
        (omniglot_mamlpp_20w5s: 2000 images per pass, inner LR 0.1), both fp32 results are two samples of the
        same chaotic spread and their distance to fp64 per tensor varies by a small factor between samples
        (measured: two builds of this engine that differ only in summation order landed at 0.4x..4.0x the
        reference's own distance over the 28 tensors) -- hence 5x, not 3x, for full-size cases.
    Conv biases are mathematically dead (BatchNorm removes them): absolute tolerance only."""
    r64 = ref64.double()
    own = float((ref32.double() - r64).abs().max())
    scale = float(r64.abs().max())
    # Mini-ImageNet 5-way 5-shot (inner LR 0.1 on 25 support images diverges): the reference's own fp32 run sits 2-60 %
    # of max-norm away from its fp64 run on most tensors and the fp32 restatement lands up to 7x that distance away
    # (measured, oracle on CPU) -- a direct comparison says nothing there; 10x keeps it as a sanity bound only.
    factor = 3.0 if not big else (10.0 if own > 0.02 * scale else 5.0)
    floor = factor * own
    if name.endswith("conv.bias") or "conv-bias" in name:
        return max(floor, 1e-5)
    return max(floor, (2e-2 if big else 2e-5) * scale + 1e-7)


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
