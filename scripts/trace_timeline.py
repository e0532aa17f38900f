"""Start-time timeline of one captured iteration from the device-side launch trace (maml_b200_trace).

    python scripts/trace_timeline.py [config] [--eval] [--first-order] [--full]

Prints, per kernel in start order, the start offset and the gap to the next start; then per kernel name the count and
the summed "time until the next kernel start" (a proxy for duration + dependency latency on the busiest chain).
"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from howtotrainyourmamlpytorch_b200 import MAMLFewShotClassifier, make_args, synthetic_batch  # noqa: E402


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    cfg = argv[0] if argv else "omniglot_mamlpp_5w1s"
    a = make_args(cfg)
    ids = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trace_kernel_ids.json")))
    names = {v: k.replace("_kernel", "") for k, v in ids.items()}
    dev = torch.device("cuda:0")
    m = MAMLFewShotClassifier(im_shape=(2, a.image_channels, a.image_height, a.image_width), device=dev, args=a)
    batch = synthetic_batch(a, iteration=0)
    epoch = 0 if "--first-order" in sys.argv else int(a.first_order_to_second_order_epoch) + 1
    run = (lambda: m.run_validation_iter(batch)) if "--eval" in sys.argv else (lambda: m.run_train_iter(batch, epoch))
    for _ in range(3):
        run()
    eng = m._engine
    eng.trace(True)
    run()
    tr = eng.trace_read()
    eng.trace(False)
    t0 = tr[0][0]
    rows = []
    for i, (t, k, tag) in enumerate(tr):
        nm = names.get(k & 0x7f, str(k)) + (":end" if k & 0x80 else "")
        nxt = tr[i + 1][0] - t if i + 1 < len(tr) else 0
        rows.append((t - t0, nm, nxt, tag))
    if "--full" in sys.argv:
        for off, nm, nxt, tag in rows:
            print("%9.2f us  %-22s +%.2f  #%d" % (off / 1e3, nm, nxt / 1e3, tag))
    agg = collections.OrderedDict()
    for off, nm, nxt, tag in rows:
        c, s = agg.get(nm, (0, 0.0))
        agg[nm] = (c + 1, s + nxt / 1e3)
    total = (tr[-1][0] - t0) / 1e3
    print("entries %d, span %.1f us" % (len(tr), total))
    for nm, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-22s n=%4d  sum-to-next %8.1f us (%.1f%%)  mean %.2f" % (nm, c, s, 100 * s / total, s / c))
    # conv_tc CTA-0 durations
    durs = []
    open_t = None
    for t, k, tag in tr:
        if k == 22:
            open_t = t
        elif k == (22 | 0x80) and open_t is not None:
            durs.append((t - open_t) / 1e3)
            open_t = None
    if durs:
        durs.sort()
        print("conv_tc CTA(0,0,0) start->end: n=%d  min %.2f  median %.2f  max %.2f us  sum %.1f" %
              (len(durs), durs[0], durs[len(durs) // 2], durs[-1], sum(durs)))


if __name__ == "__main__":
    main()
