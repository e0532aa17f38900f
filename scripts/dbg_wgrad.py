import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden
from engine_layout import theta_to_ref
from howtotrainyourmamlpytorch_b200 import MAMLFewShotClassifier
case = sys.argv[1] if len(sys.argv) > 1 else "tiny_pp"
g = load_golden(case); a = g.args
res = {}
for tag, env in (("tc", None), ("ffma", "1")):
    if env: os.environ["MAML_B200_FFMA_WGRAD"] = env
    m = MAMLFewShotClassifier(im_shape=(2, a.image_channels, a.image_height, a.image_width), device=torch.device("cuda", 0), args=a)
    m.load_state_dict(g.state()); m.meta_gradient(g.batch(0), g.iters[0][0])
    res[tag] = theta_to_ref(m._engine.debug_read("g", 0, 0, 0), a)
for l in (1, 2, 3):
    n = "classifier.layer_dict.conv%d.conv.weight" % l
    x, y = res["tc"][n], res["ffma"][n]
    print(n, "tc absmax %.3e ffma absmax %.3e" % (float(x.abs().max()), float(y.abs().max())))
    print("  tc  ", x[0, 0].flatten()[:9].numpy()); print("  ffma", y[0, 0].flatten()[:9].numpy())
    print("  tc[f=1,c=0]", x[1, 0].flatten()[:9].numpy(), " ffma", y[1, 0].flatten()[:9].numpy())
    # is tc a permutation / scaled version?
    print("  ratio stats", float((x / (y + 1e-30)).median()), "corr", float(torch.corrcoef(torch.stack([x.flatten(), y.flatten()]))[0, 1]))
