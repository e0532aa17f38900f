"""CPU: host-side logic of the drop-in surface against the golden vectors, and the C-ABI library."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ALL_CASES, ROOT, load_golden
from oracle import maml_oracle as O


def _model(g):
    from howtotrainyourmamlpytorch_b200 import MAMLFewShotClassifier
    a = g.args
    return MAMLFewShotClassifier(im_shape=(2, a.image_channels, a.image_height, a.image_width),
                                 device=torch.device("cpu"), args=a)


@pytest.mark.parametrize("case", ALL_CASES)
def test_state_dict_and_init_match_reference(case):
    g = load_golden(case)
    m = _model(g)
    sd = m.state_dict()
    ref = g.state()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]), k
    assert [n for n, p in m.named_parameters() if p.requires_grad] == list(g.grads(0).keys())
    st = O.init_state(g.args)
    for k in ref:
        assert torch.equal(st[k], ref[k]), k


@pytest.mark.parametrize("case", ALL_CASES)
def test_schedules_match_reference(case):
    g = load_golden(case)
    m = _model(g)
    for it, (epoch, _) in enumerate(g.iters):
        m.current_epoch = int(epoch)
        w = m.get_per_step_loss_importance_vector().numpy()
        assert np.array_equal(w.astype(np.float64), g.array("msl", it)), (epoch, w, g.array("msl", it))
        assert np.array_equal(w, O.msl_weights(g.args, int(epoch)))
        assert abs(m._logged_lr(int(epoch)) - g.scalar("learning_rate", it)) < 1e-12
        assert abs(m._cosine_lr(int(epoch)) - O.cosine_lr(g.args, int(epoch))) < 1e-15


def test_schedule_masks():
    g = load_golden("tiny_pp")       # MAML++: MSL for epoch < 10, S = 3
    m = _model(g)
    m.current_epoch = 0
    steps, second, mask, weights, _ = m._schedule(0, True)
    assert (steps, second, mask) == (3, True, 0b111) and abs(sum(weights) - 1.0) < 1e-6
    steps, second, mask, weights, _ = m._schedule(12, True)
    assert (steps, second, mask) == (3, True, 0b100) and weights[2] == 1.0
    steps, second, mask, weights, _ = m._schedule(0, False)
    assert (steps, second, mask) == (3, False, 0b100)
    g = load_golden("tiny_pp_first")
    m = _model(g)
    assert m._schedule(3, True)[1] is False
    sched = O.target_pass_schedule(g.args, 3, True, 3)
    assert sched == ["msl", "msl", "msl"]


def test_no_cpu_fallback():
    from howtotrainyourmamlpytorch_b200 import _native
    g = load_golden("tiny_pp")
    m = _model(g)
    with pytest.raises(_native.NativeLibraryError):
        m.run_train_iter(g.batch(0), 0)
    with pytest.raises(_native.NativeLibraryError):
        m.classifier.forward(torch.zeros(3, 3, 20, 20), num_step=0)


def test_lslr_update_rule_is_the_reference_formula():
    from howtotrainyourmamlpytorch_b200 import LSLRGradientDescentLearningRule
    rule = LSLRGradientDescentLearningRule(torch.device("cpu"), 5, True, 0.1)
    w = {"layer_dict.conv0.conv.weight": torch.randn(4, 3, 3, 3), "layer_dict.linear.bias": torch.randn(5)}
    rule.initialise(w)
    assert list(rule.names_learning_rates_dict.keys()) == ["layer_dict-conv0-conv-weight", "layer_dict-linear-bias"]
    assert rule.names_learning_rates_dict["layer_dict-linear-bias"].shape == (6,)
    gr = {k: torch.randn_like(v) for k, v in w.items()}
    out = rule.update_params(w, gr, num_step=2)
    for k in w:
        assert torch.allclose(out[k], w[k] - 0.1 * gr[k])


def test_optimizer_state_dict_roundtrip_format():
    g = load_golden("tiny_pp")
    m = _model(g)
    sd = m.optimizer.state_dict()
    ref = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in m.trainable_parameters()], lr=1e-3)
    assert set(sd["param_groups"][0].keys()) >= {"lr", "betas", "eps", "weight_decay", "amsgrad", "params"}
    assert sd["param_groups"][0]["params"] == ref.state_dict()["param_groups"][0]["params"]
    m.optimizer.step_count = 3
    m._exp_avg.normal_()
    sd = m.optimizer.state_dict()
    ref.load_state_dict(sd)            # torch's own Adam accepts it
    m2 = _model(g)
    m2.optimizer.load_state_dict(sd)
    assert m2.optimizer.step_count == 3 and torch.equal(m2._exp_avg, m._exp_avg)


def test_checkpoint_save_load(tmp_path):
    g = load_golden("tiny_maml")
    m = _model(g)
    with torch.no_grad():
        m._flat.add_(0.25)
    m.save_model(os.path.join(str(tmp_path), "train_model_latest"), {"current_iter": 7, "best_val_acc": np.float64(0.5)})
    m2 = _model(g)
    state = m2.load_model(str(tmp_path), "train_model", "latest")
    assert state["current_iter"] == 7
    assert torch.equal(m2._flat, m._flat) and m2._views_intact()


def test_loads_a_checkpoint_written_by_the_reference():
    """tests/golden/ref_ckpt_tiny_pp/train_model_latest was written by the UNMODIFIED reference's save_model after the
    recorded tiny_pp iterations (oracle/gen_golden.py --checkpoint).  load_model (reference :411-424) must restore the
    network state_dict, the Adam moments / step count and hand back the experiment state."""
    g = load_golden("tiny_pp")
    m = _model(g)
    d = os.path.join(ROOT, "tests", "golden", "ref_ckpt_tiny_pp")
    state = m.load_model(d, "train_model", "latest")
    assert state["current_iter"] == len(g.iters) and abs(state["best_val_acc"] - 0.25) < 1e-12
    post = g.post(len(g.iters) - 1)
    sd = m.state_dict()
    assert list(sd.keys()) == list(post.keys())
    for k in post:
        assert torch.equal(sd[k], post[k]), k
    assert m._views_intact()
    raw = torch.load(os.path.join(d, "train_model_latest"), map_location="cpu", weights_only=False)
    assert m.optimizer.step_count == int(float(raw["optimizer"]["state"][0]["step"])) == len(g.iters)
    for i, (name, p) in enumerate(m._trainable_param_list()):
        off, size = m._flat_slices[name]
        assert torch.equal(m._exp_avg[off:off + size].view(p.shape), raw["optimizer"]["state"][i]["exp_avg"]), name
        assert torch.equal(m._exp_avg_sq[off:off + size].view(p.shape), raw["optimizer"]["state"][i]["exp_avg_sq"]), name
    # and the other direction: what save_model writes is loadable by torch's own Adam / a plain nn.Module state_dict
    out = os.path.join(d, "..", "_roundtrip_tmp")
    try:
        m.save_model(out, {"current_iter": 3})
        again = torch.load(out, map_location="cpu", weights_only=False)
        assert list(again["network"].keys()) == list(raw["network"].keys())
        for k in raw["network"]:
            assert torch.equal(again["network"][k], raw["network"][k]), k
        for i in raw["optimizer"]["state"]:
            for f in ("exp_avg", "exp_avg_sq"):
                assert torch.equal(again["optimizer"]["state"][i][f], raw["optimizer"]["state"][i][f])
            assert float(again["optimizer"]["state"][i]["step"]) == float(raw["optimizer"]["state"][i]["step"])
        assert set(again["optimizer"]["param_groups"][0].keys()) == set(raw["optimizer"]["param_groups"][0].keys())
    finally:
        if os.path.exists(out):
            os.remove(out)


def test_c_abi_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from howtotrainyourmamlpytorch_b200 import _native
    ge.build()
    header = open(os.path.join(ROOT, "include", "maml_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(maml_b200_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert lib.maml_b200_abi_version() == _native.ABI_VERSION


def test_bench_flop_model_matches_baseline_table():
    """bench.py's algorithmic FLOPs per task (the numerator of `roofline.achieved`) against BASELINE.md section 3."""
    import importlib.util
    from howtotrainyourmamlpytorch_b200 import make_args
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    table = {"omniglot_maml_5w1s": 4.594, "omniglot_mamlpp_5w1s": 5.754, "mini_imagenet_mamlpp_5w1s": 144.60,
             "mini_imagenet_mamlpp_5w5s": 237.95, "omniglot_mamlpp_20w5s": 91.88}
    for name, gflop in table.items():
        got = bench.flops_per_task(make_args(name)) / 1e9
        assert abs(got - gflop) <= 0.005 * gflop, (name, got, gflop)


def test_batch_shape_and_label_validation():
    """The engine reads raw pointers: a batch whose shape does not match args, or labels outside [0, N), must be rejected
    on the host -- checked before any device work, so it is testable without a GPU."""
    g = load_golden("tiny_pp")
    m = _model(g)
    xs, xt, ys, yt = g.batch(0)
    with pytest.raises(ValueError):
        m._check_batch([xs[:, :, :1], xt, ys, yt])                 # wrong K
    with pytest.raises(ValueError):
        m._check_batch([xs[..., :-1], xt, ys, yt])                 # wrong W
    with pytest.raises(ValueError):
        m._check_batch([xs, xt, ys])
    assert m._check_batch([xs, xt, ys, yt]) == xs.shape[0]


def test_library_override_fails_loudly_when_the_variant_is_missing(tmp_path):
    """MAML_B200_LIB (same-box A/B of compile-time variants, scripts/build_variant.sh) selects another build of the
    engine; a path that does not exist must raise, never fall back to the default library or to a CPU path."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['MAML_B200_LIB'] = %r\n"
            "from howtotrainyourmamlpytorch_b200 import _native\n"
            "try:\n    _native.load_library()\nexcept _native.NativeLibraryError as e:\n    print('RAISED', 'missing.so' in str(e))\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "missing.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "RAISED True" in out.stdout, (out.stdout, out.stderr[-500:])
