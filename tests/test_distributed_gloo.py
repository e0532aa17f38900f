"""CPU, world_size 2 over gloo: the task-sharding scheme of the N>1 path.
Each rank evaluates its shard of tasks with the oracle, builds the result vector the way the engine's
export kernel does (gradient / loss pre-scaled by 1/B_global, running statistics pre-weighted by their
global position), ONE all_reduce(SUM) follows -- and the outcome must equal the single-process run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import maml_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_result(g, rank, world, epoch):
    from howtotrainyourmamlpytorch_b200 import sharding
    args, state = g.args, g.state()
    xs, xt, ys, yt = g.batch(0)
    B = xs.shape[0]
    assert B % world == 0
    Bl = B // world
    off, Bg = sharding.shard_of(rank, world, Bl)
    assert Bg == B
    sl = slice(off, off + Bl)
    local = (xs[sl], xt[sl], ys[sl], yt[sl])
    res = O.manual_train_iter(state, args, local, epoch, keep_intermediates=True)
    names = O.trainable_names(args)
    parts = [res["grads"][n].reshape(-1) * (Bl / float(Bg)) for n in names]
    parts.append((res["loss"] * (Bl / float(Bg))).reshape(1))
    # running statistics: weighted partial sums over the local tasks in GLOBAL order
    S = int(args.number_of_training_steps_per_iter)
    sched = O.target_pass_schedule(args, epoch, True, S)
    L = O.num_stages(args)
    F = int(args.cnn_num_filters)
    rm = torch.zeros(L, S, F, dtype=torch.float64)
    rv = torch.zeros(L, S, F, dtype=torch.float64)
    inters = [x for x in res["intermediates"] if "theta" in x]
    for x in inters:
        gidx = off + x["task"]
        for s in range(S):
            has_t = sched[s] is not None
            for l in range(L):
                blk = x["sup_f"][s]["blocks"][l]
                w = sharding.ema_weight(gidx, 0, has_t, Bg)
                rm[l, s] += w * blk["mu"].double()
                rv[l, s] += w * blk["var_unbiased"].double()
                if has_t:
                    blk = x["tgt_f"][s][0]["blocks"][l]
                    w = sharding.ema_weight(gidx, 1, has_t, Bg)
                    rm[l, s] += w * blk["mu"].double()
                    rv[l, s] += w * blk["var_unbiased"].double()
    parts += [rm.reshape(-1).float(), rv.reshape(-1).float()]
    return torch.cat([p.float() for p in parts]), names, sched


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    g = load_golden(case)
    vec, names, sched = _rank_result(g, rank, world, g.iters[0][0])
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save(vec, os.path.join(out_dir, "reduced.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["tiny_odd"])
def test_two_rank_sharding_equals_single_process(case, tmp_path):
    from howtotrainyourmamlpytorch_b200 import sharding
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    vec = torch.load(os.path.join(str(tmp_path), "reduced.pt"))
    g = load_golden(case)
    epoch = g.iters[0][0]
    full = O.manual_train_iter(g.state(), g.args, g.batch(0), epoch)
    names = O.trainable_names(g.args)
    o = 0
    for n in names:
        ref = full["grads"][n].reshape(-1)
        got = vec[o:o + ref.numel()]
        o += ref.numel()
        if "conv.bias" in n or "conv-bias" in n:
            assert float((got - ref).abs().max()) < 1e-5
        else:
            assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-8, n
    assert abs(float(vec[o]) - float(full["loss"])) <= 1e-6 * abs(float(full["loss"]))
    o += 1
    S = int(g.args.number_of_training_steps_per_iter)
    L, F = O.num_stages(g.args), int(g.args.cnn_num_filters)
    B = g.batch(0)[0].shape[0]
    sched = O.target_pass_schedule(g.args, epoch, True, S)
    mask = sum(1 << s for s in range(S) if sched[s] is not None)
    decay = sharding.decay_vector(mask, S, S, B)
    rm = vec[o:o + L * S * F].reshape(L, S, F)
    rv = vec[o + L * S * F:o + 2 * L * S * F].reshape(L, S, F)
    state = g.state()
    for l in range(L):
        _, _, _, _, rmn, rvn = O.conv_names(l)
        for s in range(S):
            new_m = decay[s] * state[rmn][s] + rm[l, s]
            new_v = decay[s] * state[rvn][s] + rv[l, s]
            assert torch.allclose(new_m, full["running"][rmn][s], rtol=1e-5, atol=1e-6)
            assert torch.allclose(new_v, full["running"][rvn][s], rtol=1e-5, atol=1e-6)
