"""Launcher shim: runs the reference's own entry point
``examples/classifier_cifar10/main_lsq.py`` UNCHANGED on this package's kernels.

    python -m cim_quantization_b200.launcher --reference-root baseline/_ref --train-batches 8 --val-batches 2

What the shim supplies, and nothing else:
* ``models._modules`` -> this package (``dropin.install()``), so ``replace_map={'Conv2d': [my_nn.Conv2dLSQCiM]}``
  (main_lsq.py:53-56) builds our modules.  ``--impl reference`` skips this step and runs the reference's own
  PyTorch-CUDA modules instead (the number a user of the reference sees today).
* stand-ins for the four third-party imports of ``examples/__init__.py:14-30`` that this image lacks
  (``plotly``, ``pytorchcv``, ``tensorboardX``, ``warmup_scheduler``) -- none of them is on the hot path.
* a synthetic CIFAR-10 (``torchvision.datasets.CIFAR10`` downloads, ``examples/__init__.py:600``; there is no
  network) and a prototxt equal to the shipped ``resnet_w3a3.prototxt`` except ``pretrained: false``, no
  ``resume``, and the epoch / batch / worker counts given on the command line.
* a wall-clock + CUDA-synchronised timer around the reference's ``train()`` so the run reports img/s.
* ``--ddp``: the reference's own multi-GPU path (``multi_gpu { multiprocessing_distributed: true }``:
  ``mp.spawn`` of ``main_worker`` per GPU, ``DistributedDataParallel``, ``DistributedSampler`` --
  examples/__init__.py:80-104, 693-716, main_lsq.py:24-59).  The spawned interpreters start from scratch, so the
  same shims are installed in them through a ``sitecustomize`` module on their PYTHONPATH
  (:func:`install_child_hooks`); every rank writes its timings to ``<work-dir>/launcher_rank<k>.json``.

The reference tree is looked up at ``--reference-root`` (default: ``$CIMQ_REFERENCE_ROOT``, ``baseline/_ref`` next
to this package, ``/root/reference``).  Nothing of it is imported by the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import runpy
import sys
import tempfile
import time
import types


def _find_reference(root: str | None) -> str:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in [root, os.environ.get("CIMQ_REFERENCE_ROOT"), os.path.join(here, "baseline", "_ref"), "/root/reference"]:
        if cand and os.path.isfile(os.path.join(cand, "examples", "classifier_cifar10", "main_lsq.py")):
            return os.path.abspath(cand)
    raise FileNotFoundError("reference tree not found (tools/make_baseline_ref.sh copies it to baseline/_ref)")


def install_stubs() -> None:
    """Inert stand-ins for the optional third-party modules examples/__init__.py imports at the top."""
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        try:
            __import__(name)
            return sys.modules[name]
        except Exception:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m

    class SummaryWriter:  # tensorboardX.SummaryWriter: the log directory must exist (prototxt dump, checkpoints)
        def __init__(self, logdir=None, *a, **k):
            if logdir:
                os.makedirs(logdir, exist_ok=True)

        def __getattr__(self, name):
            return lambda *a, **k: None

    class GradualWarmupScheduler:  # only constructed when the prototxt has a `warmup` block
        def __init__(self, *a, **k):
            raise NotImplementedError("warmup_scheduler is not installed")

    def ptcv_get_model(*a, **k):
        raise NotImplementedError("pytorchcv is not installed")

    plotly = mod("plotly")
    go = mod("plotly.graph_objects")
    if not hasattr(plotly, "graph_objects"):
        plotly.graph_objects = go
    ptcv = mod("pytorchcv")
    mp_ = mod("pytorchcv.model_provider", get_model=ptcv_get_model)
    if not hasattr(ptcv, "model_provider"):
        ptcv.model_provider = mp_
    mod("tensorboardX", SummaryWriter=SummaryWriter)
    mod("warmup_scheduler", GradualWarmupScheduler=GradualWarmupScheduler)


def install_synthetic_cifar10(train_images: int, val_images: int, seed: int = 0) -> None:
    """torchvision.datasets.CIFAR10 -> deterministic random 32x32 RGB images with the same interface."""
    import numpy as np
    import torchvision
    from PIL import Image

    class SyntheticCIFAR10:
        classes = [str(i) for i in range(10)]

        def __init__(self, root=None, train=True, transform=None, target_transform=None, download=False):
            n = train_images if train else val_images
            rng = np.random.RandomState(seed + (0 if train else 1))
            self.data = rng.randint(0, 256, size=(n, 32, 32, 3), dtype=np.uint8)
            self.targets = rng.randint(0, 10, size=(n,)).tolist()
            self.transform, self.target_transform = transform, target_transform

        def __len__(self):
            return len(self.targets)

        def __getitem__(self, i):
            img, t = Image.fromarray(self.data[i]), self.targets[i]
            if self.transform is not None:
                img = self.transform(img)
            if self.target_transform is not None:
                t = self.target_transform(t)
            return img, t

    torchvision.datasets.CIFAR10 = SyntheticCIFAR10


def write_prototxt(ref_root: str, path: str, epochs: int, batch_size: int, workers: int, overrides: dict) -> None:
    """The shipped prototxt with the offline-only fields changed (text-level edit, parsed by the reference)."""
    src = os.path.join(ref_root, "examples", "classifier_cifar10", "prototxt", "resnet_w3a3.prototxt")
    keep = []
    fixed = {"pretrained": "false", "epochs": str(epochs), "batch_size": str(batch_size), "workers": str(workers),
             "print_freq": "4", "log_name": '"launcher"'}
    fixed.update({k: str(v) for k, v in overrides.items()})
    seen = set()
    for line in open(src):
        key = line.split(":")[0].strip()
        if key == "resume":
            continue  # a checkpoint path on the author's machine
        if key in fixed:
            keep.append(f"{key}: {fixed[key]}\n")
            seen.add(key)
        else:
            keep.append(line)
    for k, v in fixed.items():
        if k not in seen:
            keep.insert(0, f"{k}: {v}\n")
    with open(path, "w") as f:
        f.writelines(keep)


CHILD_ENV = "CIMQ_LAUNCHER_CHILD"


def _timed_train_patch(examples, stats, sink=None):
    import torch
    ref_train = examples.train

    def timed_train(train_loader, model, criterion, optimizer, epoch, args, writer):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ref_train(train_loader, model, criterion, optimizer, epoch, args, writer)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sampler = getattr(train_loader, "sampler", None)
        n = len(sampler) if sampler is not None and hasattr(sampler, "__len__") else len(train_loader.dataset)
        stats["epochs"].append({"epoch": epoch, "seconds": dt, "images": n, "img_per_s": n / dt,
                                "model_class": type(model).__name__})
        if sink is not None:
            sink(args)
        return out

    examples.train = timed_train


def install_child_hooks() -> None:
    """Called from the generated ``sitecustomize`` in interpreters spawned by the reference's ``mp.spawn``: the same
    shims as the parent (stubs, synthetic CIFAR-10, ``models._modules`` -> this package, timed ``train``)."""
    cfg = os.environ.get(CHILD_ENV)
    if not cfg:
        return
    c = json.loads(cfg)
    install_stubs()
    install_synthetic_cifar10(c["train_images"], c["val_images"])
    if c["ref_root"] not in sys.path:
        sys.path.insert(0, c["ref_root"])
    if c["impl"] == "ours":
        from . import dropin
        dropin.install()
    import examples
    examples.get_freer_gpu = lambda: 0
    # Reference defect worked around, in the spawned ranks only: get_summary_writer sets ``args.log_name`` on the first
    # rank of a node alone (examples/__init__.py:466-476) but main_lsq.py:77 reads it on EVERY rank, so the reference's
    # own DDP path dies with AttributeError on ranks > 0.  The other ranks get a scratch directory name of their own.
    ref_writer = examples.get_summary_writer

    def get_summary_writer(args, ngpus_per_node, model):
        w = ref_writer(args, ngpus_per_node, model)
        if not hasattr(args, "log_name"):
            args.log_name = "logger/_rank%d" % int(getattr(args, "gpu", 0) or 0)
            os.makedirs(args.log_name, exist_ok=True)
        return w

    examples.get_summary_writer = get_summary_writer
    stats = {"epochs": []}

    def sink(args):
        conv_cls = sys.modules["models._modules"].Conv2dLSQCiM  # in the parent; the ranks report theirs
        stats["conv_class"] = conv_cls.__module__ + "." + conv_cls.__name__
        stats["rank"] = int(getattr(args, "gpu", 0) or 0)
        with open(os.path.join(c["work"], f"launcher_rank{stats['rank']}.json"), "w") as f:
            json.dump(stats, f)

    _timed_train_patch(examples, stats, sink)


def run(argv=None) -> dict:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference-root", default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--train-batches", type=int, default=8, help="synthetic training set = this many batches")
    ap.add_argument("--val-batches", type=int, default=2)
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE", help="extra prototxt overrides")
    ap.add_argument("--work-dir", default=None, help="where the reference writes ./logger/... (default: a temp dir)")
    ap.add_argument("--ddp", action="store_true",
                    help="the reference's own DDP path: multi_gpu { multiprocessing_distributed: true }, one spawned "
                         "process per visible GPU; --batch-size is then the GLOBAL batch (examples/__init__.py:707)")
    ap.add_argument("--ddp-port", type=int, default=23456)
    a = ap.parse_args(argv)

    import torch
    ref_root = _find_reference(a.reference_root)
    install_stubs()
    install_synthetic_cifar10(a.train_batches * a.batch_size, a.val_batches * a.batch_size)
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    if a.impl == "ours":
        from . import dropin
        dropin.install()  # before `import examples`: `import models._modules as my_nn` must resolve to this package
    import examples  # the reference's harness (examples/__init__.py)

    examples.get_freer_gpu = lambda: 0  # nvidia-smi text parsing (examples/__init__.py:117-130) is brittle
    stats = {"epochs": []}
    _timed_train_patch(examples, stats)
    work = a.work_dir or tempfile.mkdtemp(prefix="cimq_launcher_")
    os.makedirs(work, exist_ok=True)
    hp = os.path.join(work, "resnet_w3a3_offline.prototxt")
    overrides = dict(kv.split("=", 1) for kv in a.set)
    write_prototxt(ref_root, hp, a.epochs, a.batch_size, a.workers, overrides)
    if a.ddp:
        with open(hp, "a") as f:  # the reference's own switch for one process per GPU + DistributedDataParallel
            f.write('multi_gpu {\n  world_size: 1\n  rank: 0\n  dist_url: "tcp://127.0.0.1:%d"\n  dist_backend: "nccl"\n'
                    '  multiprocessing_distributed: true\n}\n' % a.ddp_port)
        hook_dir = os.path.join(work, "_child_hooks")
        os.makedirs(hook_dir, exist_ok=True)
        with open(os.path.join(hook_dir, "sitecustomize.py"), "w") as f:
            f.write("import os\nif os.environ.get(%r):\n    from cim_quantization_b200.launcher import install_child_hooks\n"
                    "    install_child_hooks()\n" % CHILD_ENV)
        pkg_parent = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.environ["PYTHONPATH"] = os.pathsep.join([hook_dir, pkg_parent, os.environ.get("PYTHONPATH", "")])
        os.environ[CHILD_ENV] = json.dumps({"ref_root": ref_root, "impl": a.impl, "work": work,
                                            "train_images": a.train_batches * a.batch_size,
                                            "val_images": a.val_batches * a.batch_size})
    cwd, argv0 = os.getcwd(), sys.argv
    os.chdir(work)  # the reference writes ./logger/... and copies its sources there
    sys.argv = ["main_lsq.py", "--hp", hp]
    try:
        runpy.run_path(os.path.join(ref_root, "examples", "classifier_cifar10", "main_lsq.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = argv0
    conv_cls = sys.modules["models._modules"].Conv2dLSQCiM
    ckpts = []
    for root_, _, files in os.walk(os.path.join(work, "logger")):
        ckpts += [os.path.join(root_, f) for f in files if f.endswith(".pth.tar")]
    stats["checkpoints"] = sorted(ckpts, key=os.path.getmtime)  # written by the reference's save_checkpoint
    stats["work_dir"] = work
    if a.ddp:
        os.environ.pop(CHILD_ENV, None)
        ranks = []
        for f in sorted(os.listdir(work)):
            if f.startswith("launcher_rank") and f.endswith(".json"):
                ranks.append(json.load(open(os.path.join(work, f))))
        stats["ranks"] = ranks
        if ranks:  # epoch time = the slowest rank's; images = all ranks' shards
            ne = min(len(r["epochs"]) for r in ranks)
            stats["epochs"] = [{"epoch": e, "seconds": max(r["epochs"][e]["seconds"] for r in ranks),
                                "images": sum(r["epochs"][e]["images"] for r in ranks),
                                "img_per_s": sum(r["epochs"][e]["images"] for r in ranks) /
                                max(r["epochs"][e]["seconds"] for r in ranks),
                                "model_class": ranks[0]["epochs"][e].get("model_class")} for e in range(ne)]
            stats["world_size"] = len(ranks)
    stats.update({"impl": a.impl, "conv_class": conv_cls.__module__ + "." + conv_cls.__name__, "batch_size": a.batch_size,
                  "reference_root": ref_root})
    print("LAUNCHER_RESULT " + json.dumps(stats))
    return stats


if __name__ == "__main__":
    run()
