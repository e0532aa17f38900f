#!/usr/bin/env python
"""Benchmark of the MAML / MAML++ hot path (BASELINE.json metric: meta-tasks/sec, 5-way, 5 inner steps).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--impl ours|reference]

One "step" = one ``run_train_iter`` over one meta-batch of synthetic episodes: inner-loop unroll for every
task, second-order meta-gradient, (all-reduce over ranks), clamp + Adam, running-stat EMA.
  value      whole-job tasks/s with the episode tensors already resident in HBM, per-step CUDA-event timing,
             L2 flushed between steps, max over ranks;
  e2e        the same metric through the public API ``MAMLFewShotClassifier.run_train_iter`` with HOST
             tensors: pinned H2D of the episodes and D2H of loss / accuracy / logits inside the timed region;
  roofline   the dominant kernel class (3x3 implicit-GEMM convolutions): algorithmic conv FLOPs per launch
              / mean launch duration from CUDA events on the launching stream, against the
             measured tensor peak (MEASURED_PEAKS.json bf16 / 2 = TF32, / 3 for the fp32-faithful 3xTF32 split);
  cpu_baseline  the reference's CPU path restated (oracle "port": same torch.nn.functional ops + autograd as
             the reference -- the reference itself is Python and cannot travel to the GPU box), timed on the
             host cores on a bounded sample of the same workload.
``--impl reference`` prints the CPU arm as its own line (rank 0 only under torchrun).
Weak scaling: every rank holds ``batch_size`` tasks (tasks are sharded over GPUs, one all-reduce of the flat
meta-gradient per iteration); the global meta-batch is N x batch_size.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "meta-tasks/sec (5-way, 5 inner steps, second order)"
DEFAULT_CONFIG = "omniglot_mamlpp_5w1s"          # BASELINE.json configs[1]: the 1xB200 headline workload


def workload_desc(name, args, n_gpus):
    return {
        "workload": "%s: %d-way %d-shot, %d target/class, %dx%dx%d, %d filters, %d inner steps, meta-batch %d per GPU"
                    % (name, args.num_classes_per_set, args.num_samples_per_class, args.num_target_samples,
                       args.image_height, args.image_width, args.image_channels, args.cnn_num_filters,
                       args.number_of_training_steps_per_iter, args.batch_size),
        "config": name, "tasks_per_gpu": int(args.batch_size), "global_meta_batch": int(args.batch_size) * n_gpus,
        "second_order": bool(args.second_order), "multi_step_loss": bool(args.use_multi_step_loss_optimization),
        "per_step_bn": bool(args.per_step_bn_statistics), "parallelism": "task-sharded dp%d" % n_gpus,
        "l2": "flushed between steps (256 MiB memset outside the per-step event pair); 8 distinct episode batches cycled",
        "inputs": "bernoulli(0.93) 28x28x1" if args.image_channels == 1 else "normal(0,1) 84x84x3",
    }


def flops_per_task(args):
    """Algorithmic conv FLOPs per task."""
    h, w, c = args.image_height, args.image_width, args.image_channels
    F = args.cnn_num_filters
    fl = []
    for _ in range(args.num_stages):
        fl.append(2.0 * h * w * F * c * 9)
        h, w, c = h // 2, w // 2, F
    n_s = args.num_classes_per_set * args.num_samples_per_class
    n_t = args.num_classes_per_set * args.num_target_samples
    S = args.number_of_training_steps_per_iter
    sup2 = 4 * fl[0] + 9 * sum(fl[1:])
    sup1 = 2 * fl[0] + 3 * sum(fl[1:])
    tgt = 2 * fl[0] + 3 * sum(fl[1:])
    n_tp = S if args.use_multi_step_loss_optimization else 1
    return S * n_s * (sup2 if args.second_order else sup1) + n_tp * n_t * tgt


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def _cpu_port_iteration_times(args, iters, warmup, threads):
    import torch
    from oracle import maml_oracle as O
    torch.set_num_threads(threads)
    state = O.init_state(args)
    names = O.trainable_names(args)
    m = {n: torch.zeros_like(state[n]) for n in names}
    v = {n: torch.zeros_like(state[n]) for n in names}
    step, times = 0, []
    for it in range(warmup + iters):
        batch = O.synthetic_batch(args, iteration=it)
        t0 = time.perf_counter()
        res = O.autograd_train_iter(state, args, batch, 0)
        clamp = [n for n in names if n.startswith("classifier.")] if "imagenet" in args.dataset_name else None
        newp, m, v, step = O.adam_step({n: state[n] for n in names}, res["grads"], m, v, step, O.cosine_lr(args, 0), clamp=clamp)
        state.update(newp)
        state.update(res["running"])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


def cpu_port_tasks_per_sec(args, iters, warmup, threads=None):
    """The reference's CPU path (restated: same torch.nn.functional ops + autograd.grad(create_graph) + one reverse
    sweep as reference few_shot_learning_system.py:170-263,325-336) on the host cores.  The thread count is tuned
    (1 probe iteration each over 8/16/32/64/all cores -- these ops are small, more threads is not faster) and the
    best is used; ``cores`` reports the threads actually used.  Returns (tasks/s, cores, sample, times)."""
    ncpu = os.cpu_count() or 1
    if threads is None:
        cands = sorted(set(c for c in (8, 16, 32, 64, ncpu) if c <= ncpu)) or [ncpu]
        best, best_t = cands[0], None
        for c in cands:
            t = _cpu_port_iteration_times(args, 1, 1, c)[0]
            if best_t is None or t < best_t:
                best, best_t = c, t
            if t > 4.0 * best_t:
                break
        threads = best
    times = sorted(_cpu_port_iteration_times(args, iters, warmup, threads))
    med = times[len(times) // 2]
    sample = "%d timed iterations of %d tasks (median), %d warm-up, %d of %d host threads (tuned)" % (
        iters, args.batch_size, warmup, threads, ncpu)
    return args.batch_size / med, threads, sample, times


def torch_gpu_port_tasks_per_sec(args, dev, iters=3, warmup=1):
    """SURVEY.md section 8d "library kernels to beat": the same call-for-call restatement of the reference
    (torch.nn.functional convs / batch_norm / max_pool2d + autograd.grad(create_graph) + one reverse sweep), but on the
    GPU through PyTorch's own CUDA kernels (cuDNN / ATen), strict fp32 (TF32 off).  This is what the reference does when
    it sees a GPU (few_shot_learning_system.py:73-81).  A baseline beside the line, never the thing measured."""
    import torch
    from oracle import maml_oracle as O
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        state = {k: v.to(dev) for k, v in O.init_state(args).items()}
        names = O.trainable_names(args)
        m = {n: torch.zeros_like(state[n]) for n in names}
        v = {n: torch.zeros_like(state[n]) for n in names}
        step, times = 0, []
        for it in range(warmup + iters):
            batch = tuple(t.to(dev) for t in O.synthetic_batch(args, iteration=it))
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            res = O.autograd_train_iter(state, args, batch, 0)
            clamp = [n for n in names if n.startswith("classifier.")] if "imagenet" in args.dataset_name else None
            newp, m, v, step = O.adam_step({n: state[n] for n in names}, res["grads"], m, v, step, O.cosine_lr(args, 0), clamp=clamp)
            state.update(newp)
            state.update(res["running"])
            torch.cuda.synchronize(dev)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        times.sort()
        med = times[len(times) // 2]
        return {"value": args.batch_size / med, "unit": "tasks/s", "ms_per_step": 1e3 * med,
                "kind": "port on PyTorch CUDA library kernels (cuDNN / ATen eager autograd), fp32, inputs resident",
                "sample": "%d timed iterations of %d tasks (median), %d warm-up" % (iters, args.batch_size, warmup)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def run_reference_arm(cli, args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    steps = max(1, cli.steps)
    warm = max(1, min(cli.warmup, 2))
    # bounded: one step = one meta-batch on the CPU (about 1 s for the Omniglot workload); cap the total work
    t_probe0 = time.perf_counter()
    tps, cores, sample, times = cpu_port_tasks_per_sec(args, iters=min(steps, 20), warmup=warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tasks/s", "n_gpus": world, "steps": len(times),
        "warmup": warm, "ms_per_step": 1e3 * sorted(times)[len(times) // 2], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": workload_desc(cli.config, args, 1),
        "cpu_baseline": {"value": tps, "unit": "tasks/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = PyTorch eager CPU path restated call for call (oracle/maml_oracle.py autograd_train_iter); "
                "the Python reference itself cannot travel to the GPU box; wall %.1f s" % (time.perf_counter() - t_probe0),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=str, default=DEFAULT_CONFIG)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-size", type=int, default=None, help="tasks per GPU (default: the config's batch_size)")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: do not flush L2 between timed steps")
    ap.add_argument("--sync-each-step", action="store_true", help="diagnostic: synchronize after every timed step")
    cli = ap.parse_args()

    import torch
    from howtotrainyourmamlpytorch_b200 import make_args

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    over = {}
    if cli.batch_size:
        over["batch_size"] = cli.batch_size
    args = make_args(cli.config, **over)

    if cli.impl == "reference":
        run_reference_arm(cli, args, rank, world)
        return

    import torch.distributed as dist
    from howtotrainyourmamlpytorch_b200 import MAMLFewShotClassifier, synthetic_batch, _native
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    W, K = max(3, cli.warmup), max(1, cli.steps)

    model = MAMLFewShotClassifier(im_shape=(2, args.image_channels, args.image_height, args.image_width), device=dev, args=args)
    B = int(args.batch_size)
    n_pool = 8
    host_batches = [synthetic_batch(args, iteration=1000 * rank + i) for i in range(n_pool)]
    dev_batches = [(hb[0].to(dev), hb[1].to(dev), hb[2].long().to(dev), hb[3].long().to(dev)) for hb in host_batches]
    pinned_batches = [tuple(t.pin_memory() for t in hb) for hb in host_batches]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        # inputs already resident in HBM; no host sync inside
        model._current_lr = model._cosine_lr(0)
        return model._run(dev_batches[i % n_pool], 0, training_phase=True, apply_update=True)

    # ---------------- device-resident throughput (value)
    # warm-up: at least W steps and at least one pass over every distinct episode batch (the engine captures one
    # CUDA graph per distinct set of buffer addresses; captures belong to warm-up, not to the timed region)
    for i in range(max(W, n_pool)):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    wall0 = time.perf_counter()
    for i in range(K):
        if not cli.no_flush:
            flush.zero_()
        ev[i][0].record()
        device_step(W + i)
        ev[i][1].record()
        if cli.sync_each_step:
            torch.cuda.synchronize()
    barrier()
    wall_dev = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    launches_per_step = model._engine.last_launch_count() + 1 + (1 if args.per_step_bn_statistics else 0)

    # ---------------- end to end through the public API with host buffers (e2e)
    for i in range(max(3, W)):
        model.run_train_iter(pinned_batches[i % n_pool], 0)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        losses, preds = model.run_train_iter(pinned_batches[(W + i) % n_pool], 0)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    h2d = sum(t.numel() * (8 if j >= 2 else 4) for j, t in enumerate(host_batches[0]))   # images fp32, labels int64 on the device
    n_t = args.num_classes_per_set * args.num_target_samples
    d2h = 2 * 4 + B * n_t * args.num_classes_per_set * 4

    # ---------------- roofline leg: per-launch CUDA events by kernel class (separate, un-timed pass)
    eng = model._engine
    eng.profile(True)
    prof_steps = 3
    for i in range(prof_steps):
        device_step(i)
    prof = eng.profile_read()
    eng.profile(False)

    # max over ranks
    t_dev = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t_dev[0]), float(t_dev[1])

    if rank == 0:
        peaks, peak_src = measured_peaks()
        tasks_total = B * world * K
        value = tasks_total / (total_ms * 1e-3)
        e2e_value = tasks_total / (e2e_ms * 1e-3)
        fpt = flops_per_task(args)
        conv_ms, conv_fl, conv_n = prof["conv_igemm"]
        wg_ms, wg_fl, wg_n = prof["wgrad"]
        tot_prof_ms = sum(v[0] for v in prof.values())
        # dominant kernel class = the implicit-GEMM convolutions (forward / tangent / dgrad) + wgrad
        dom_ms, dom_fl, dom_n = conv_ms + wg_ms, conv_fl + wg_fl, conv_n + wg_n
        tf32_peak = peaks["bf16_tflops"] / 2.0            # dense TF32 = half of dense bf16 (measured burst)
        peak_3x = tf32_peak / 3.0                         # fp32-faithful 3xTF32 operand split
        achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r1.json")))["dominant_kernel_traffic_bytes_per_launch"]
        except Exception:
            pass
        roofline = {
            "bound": "tensor", "kernel": "3x3 implicit-GEMM conv: conv_tc_kernel (tcgen05, 3xTF32, cluster split-K) + wgrad_row_kernel (FFMA)",
            "achieved": achieved, "peak": peak_3x, "unit": "TFLOP/s", "frac": achieved / peak_3x,
            "traffic": traffic,
            "traffic_note": "dram__bytes_read+write of one block-1 conv_tc_kernel launch, grid (10,8,1) (ncu --set full, cold cache; profiles/ncu_summary_r1.json)",
            "peak_source": peak_src + ": bf16_tflops %.1f / 2 (tf32) / 3 (3xTF32 split)" % peaks["bf16_tflops"],
            "launches_profiled": int(dom_n), "mean_launch_us": 1e3 * dom_ms / max(dom_n, 1),
            "share_of_step": dom_ms / tot_prof_ms if tot_prof_ms > 0 else None,
            "whole_iteration": {"alg_tflops": value * fpt / 1e12, "frac_of_peak": value * fpt / 1e12 / (peak_3x * world)},
            "by_class_ms_per_step": {k: v[0] / prof_steps for k, v in prof.items()},
        }
        line = {
            "metric": METRIC, "value": value, "unit": "tasks/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": workload_desc(cli.config, args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "tasks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches_per_step * K),
            "roofline": roofline,
            "gflop_per_task": fpt / 1e9,
            "wall_s_device_loop": wall_dev,
            "step_ms": {"min": min(step_ms), "median": sorted(step_ms)[len(step_ms) // 2], "max": max(step_ms)},
            "workspace_mib": eng.workspace_bytes / 2 ** 20,
            "last_loss": float(losses["loss"]),
        }
        if not cli.no_cpu_baseline and world == 1:
            try:
                line["torch_gpu_baseline"] = torch_gpu_port_tasks_per_sec(args, dev)
            except Exception as exc:      # a baseline must never take the measurement down
                line["torch_gpu_baseline"] = {"unavailable": repr(exc)[:200]}
            tps, cores, sample, _ = cpu_port_tasks_per_sec(args, iters=8, warmup=2)
            line["cpu_baseline"] = {"value": tps, "unit": "tasks/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
