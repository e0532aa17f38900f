"""Stage the UNMODIFIED reference under ``baseline/_ref/`` (git-ignored, shipped to the GPU box by gpurun).

The reference (AntreasAntoniou/HowToTrainYourMAMLPytorch) is a pure-Python program without packaging metadata
(no setup.py / pyproject.toml), so ``pip install --target baseline/_ref /root/reference`` has nothing to build; what an
install would amount to is making its modules importable from one directory, which is what this script does: it copies
the reference's own ``.py`` modules (and its ``utils`` package) byte for byte.  Nothing is edited, nothing of it is
committed -- ``baseline/_ref/`` is in ``.gitignore`` -- and ``MANIFEST.json`` records the sha256 of every staged file.

  python baseline/stage_reference.py [/root/reference]

``bench.py --impl reference`` and the ``cpu_baseline`` / ``torch_gpu_baseline`` legs import the staged copy through
``baseline/run_reference.py``; ``/root/reference`` itself does not exist on the GPU box.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
MODULES = ["few_shot_learning_system.py", "meta_neural_network_architectures.py", "inner_loop_optimizers.py",
           "experiment_builder.py", "data.py", "train_maml_system.py",
           os.path.join("utils", "__init__.py"), os.path.join("utils", "parser_utils.py"),
           os.path.join("utils", "storage.py"), os.path.join("utils", "dataset_tools.py")]


def stage(src="/root/reference"):
    if not os.path.isdir(src):
        return None
    manifest = {}
    for rel in MODULES:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    cfg_src, cfg_dst = os.path.join(src, "experiment_config"), os.path.join(DEST, "experiment_config")
    if os.path.isdir(cfg_src):
        shutil.copytree(cfg_src, cfg_dst, dirs_exist_ok=True)
    sub = os.path.join(src, ".SUBMODULES.json")
    commit = None
    if os.path.exists(sub):
        try:
            commit = json.load(open(sub)).get("commit")
        except Exception:
            commit = None
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "commit": commit, "sha256": manifest}, fh, indent=1)
    return DEST


if __name__ == "__main__":
    out = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print(out if out else "reference sources not found; nothing staged")
