"""One-process validation of a scheduling-only engine switch: (1) the meta-gradient with the switch off / on agrees on tiny and
full-size golden cases, (2) in-process timing of both settings on the headline workload.

    python scripts/check_zstage.py [ENV_VAR]          (default MAML_B200_TC_ZSTAGE)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
VAR = sys.argv[1] if len(sys.argv) > 1 else "MAML_B200_TC_ZSTAGE"


def main():
    import torch
    import bench
    from conftest import load_golden
    from engine_layout import rel_err
    from howtotrainyourmamlpytorch_b200 import MAMLFewShotClassifier, make_args
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    t0 = time.time()

    def grads(case, val):
        os.environ[VAR] = val
        g = load_golden(case)
        a = g.args
        m = MAMLFewShotClassifier(im_shape=(2, a.image_channels, a.image_height, a.image_width), device=dev, args=a)
        m.load_state_dict(g.state())
        l, p, gr = m.meta_gradient(g.batch(0), g.iters[0][0])
        return float(l["loss"]), {k: v.detach().clone() for k, v in gr.items()}

    worst = 0.0
    for case in ("tiny_pp", "tiny_bern", "omniglot_mamlpp_5w1s", "mini_imagenet_mamlpp_5w1s", "omniglot_mamlpp_20w5s"):
        try:
            l0, g0 = grads(case, "0")
            l1, g1 = grads(case, "1")
            e = max(rel_err(g1[n], g0[n]) for n in g0 if "conv.bias" not in n and "conv-bias" not in n)
            worst = max(worst, e)
            print("%-28s loss %.7f / %.7f  max rel err %.2e  %s" % (case, l0, l1, e, "OK" if e <= 2e-5 and abs(l0 - l1) <= 1e-6 * abs(l0) else "MISMATCH"), flush=True)
        except Exception as exc:
            print("%-28s ERROR %r" % (case, exc), flush=True)
            worst = float("inf")
    print("equivalence worst rel err %.2e (%.1f s)" % (worst, time.time() - t0), flush=True)

    args = make_args("omniglot_mamlpp_5w1s")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    res = {"0": [], "1": []}
    for rnd in range(2):
        for val in ("1", "0"):
            os.environ[VAR] = val
            model = MAMLFewShotClassifier(im_shape=(2, args.image_channels, args.image_height, args.image_width), device=dev, args=args)
            r = bench.measure_device_loop(model, args, dev, 0, 1, 20, 4, flush)
            ms = sorted(r["step_ms"])
            res[val].append(sum(ms) / len(ms))
            del model, r
    print("headline ms/step  %s=1: %s   %s=0: %s" % (VAR, ["%.4f" % x for x in res["1"]], VAR, ["%.4f" % x for x in res["0"]]), flush=True)


if __name__ == "__main__":
    main()
