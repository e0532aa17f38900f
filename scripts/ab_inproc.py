"""In-process A/B of engine switches (one python start-up for many variants): every variant builds a fresh model with
its environment variables set (the engine reads them at handle creation), runs bench.py's device-resident loop and
prints mean / median ms per step.

    python scripts/ab_inproc.py [--config NAME] [--batch-size B] [--steps K] [--rounds R] "VAR=1 VAR2=0" "VAR=2" ...

The empty string "" is the default build.  Variants are interleaved R times (round-robin) so that slow drift of the box
does not masquerade as a difference; the table reports the minimum of the per-round means.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="omniglot_mamlpp_5w1s")
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("variants", nargs="*")
    cli = ap.parse_args()
    import torch
    import bench
    from howtotrainyourmamlpytorch_b200 import MAMLFewShotClassifier, make_args

    over = {"batch_size": cli.batch_size} if cli.batch_size else {}
    args = make_args(cli.config, **over)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    variants = cli.variants or [""]
    touched = set()
    for v in variants:
        for kv in v.split():
            touched.add(kv.split("=", 1)[0])
    res = {v: [] for v in variants}
    for rnd in range(cli.rounds):
        for v in variants:
            for k in touched:
                os.environ.pop(k, None)
            for kv in v.split():
                k, val = kv.split("=", 1)
                os.environ[k] = val
            try:
                model = MAMLFewShotClassifier(im_shape=(2, args.image_channels, args.image_height, args.image_width), device=dev, args=args)
                r = bench.measure_device_loop(model, args, dev, 0, 1, cli.steps, cli.warmup, flush)
                ms = sorted(r["step_ms"])
                res[v].append({"mean": sum(ms) / len(ms), "median": ms[len(ms) // 2], "min": ms[0]})
                del model, r
                torch.cuda.empty_cache()
            except Exception as exc:      # a broken variant must not take the others down
                res[v].append({"error": repr(exc)[:200]})
            print("round %d  %-60s %s" % (rnd, v or "(default)", res[v][-1]), flush=True)
    print("\n%-60s %10s %10s" % ("variant", "best mean", "best median"))
    summary = {}
    for v in variants:
        ok = [x for x in res[v] if "mean" in x]
        if ok:
            summary[v or "(default)"] = {"mean_ms": min(x["mean"] for x in ok), "median_ms": min(x["median"] for x in ok)}
            print("%-60s %10.4f %10.4f" % (v or "(default)", summary[v or "(default)"]["mean_ms"], summary[v or "(default)"]["median_ms"]))
        else:
            summary[v or "(default)"] = res[v]
            print("%-60s FAILED %s" % (v or "(default)", res[v]))
    if cli.out:
        json.dump({"config": cli.config, "batch_size": cli.batch_size, "steps": cli.steps, "rounds": cli.rounds, "results": summary, "raw": res},
                  open(cli.out, "w"), indent=1)


if __name__ == "__main__":
    main()
