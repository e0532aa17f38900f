"""Diagnostic: where does one iteration go?  Times (CUDA events, graph replay) of
  eval (phase A support chain + 1 target forward), first-order training (phase A + target backward + alpha-bar),
  full second-order training, for a BASELINE config."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from howtotrainyourmamlpytorch_b200 import MAMLFewShotClassifier, make_args, synthetic_batch

name = sys.argv[1] if len(sys.argv) > 1 else "omniglot_mamlpp_5w1s"
dev = torch.device("cuda", 0)
for label, over in (("second-order", {}), ("first-order", {"second_order": False})):
    a = make_args(name, **over)
    m = MAMLFewShotClassifier(im_shape=(2, a.image_channels, a.image_height, a.image_width), device=dev, args=a)
    b = synthetic_batch(a, 0)
    db = (b[0].to(dev), b[1].to(dev), b[2].long().to(dev), b[3].long().to(dev))
    for mode in ("train", "eval"):
        f = (lambda: m._run(db, 0, True, True)) if mode == "train" else (lambda: m._run(db, 0, False, False))
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        print("%-28s %-13s %-6s %.3f ms/iter  (%d launches)" % (name, label, mode, e0.elapsed_time(e1) / 20, m._engine.last_launch_count()))
