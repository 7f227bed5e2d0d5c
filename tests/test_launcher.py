"""Launcher shim: the pieces that let the reference's main_lsq.py run unchanged.  The CPU tests
cover the stand-in modules, the synthetic CIFAR-10 and the offline prototxt; the GPU test runs the reference's
own entry point end to end on our kernels (needs the baseline/_ref copy made by tools/make_baseline_ref.sh)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
has_ref = os.path.isfile(os.path.join(REF, "examples", "classifier_cifar10", "main_lsq.py"))


def test_stubs_and_synthetic_dataset():
    from cim_quantization_b200 import launcher
    launcher.install_stubs()
    import tensorboardX
    import warmup_scheduler  # noqa: F401
    from pytorchcv.model_provider import get_model  # noqa: F401
    import plotly.graph_objects  # noqa: F401
    w = tensorboardX.SummaryWriter(None)
    w.add_scalar("a", 1.0, 0)  # every method is a no-op
    import torchvision
    saved = torchvision.datasets.CIFAR10
    try:
        launcher.install_synthetic_cifar10(12, 5)
        tr = torchvision.datasets.CIFAR10(root="x", train=True, download=True,
                                          transform=torchvision.transforms.ToTensor())
        va = torchvision.datasets.CIFAR10(root="x", train=False, download=True)
        assert len(tr) == 12 and len(va) == 5
        img, t = tr[3]
        assert tuple(img.shape) == (3, 32, 32) and 0 <= t < 10
        img2, t2 = torchvision.datasets.CIFAR10(root="x", train=True)[3]
        assert t2 == t  # deterministic
    finally:
        torchvision.datasets.CIFAR10 = saved


@pytest.mark.skipif(not has_ref, reason="baseline/_ref not present (tools/make_baseline_ref.sh)")
def test_offline_prototxt_parses_with_the_reference_schema(tmp_path):
    from cim_quantization_b200 import launcher
    path = str(tmp_path / "hp.prototxt")
    launcher.write_prototxt(REF, path, epochs=2, batch_size=64, workers=0, overrides={"xbar": 64})
    sys.path.insert(0, REF)
    try:
        import google.protobuf.text_format as tf
        from proto import efficient_pytorch_pb2 as eppb
        hp = eppb.HyperParam()
        tf.Merge(open(path).read(), hp)
    finally:
        sys.path.remove(REF)
    assert hp.pretrained is False and hp.epochs == 2 and hp.batch_size == 64 and hp.xbar == 64
    assert not hp.HasField("resume")
    assert hp.nbits_w == 3 and hp.nbits_a == 3 and abs(hp.adcbits - 1.5) < 1e-6  # untouched fields survive


@pytest.mark.gpu
@pytest.mark.skipif(not has_ref, reason="baseline/_ref not present (tools/make_baseline_ref.sh)")
def test_reference_main_lsq_runs_unchanged_on_our_kernels():
    import subprocess
    r = subprocess.run([sys.executable, "-m", "cim_quantization_b200.launcher", "--impl", "ours", "--train-batches", "3",
                        "--val-batches", "1", "--batch-size", "64", "--workers", "0"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("LAUNCHER_RESULT ")][-1]
    import json
    res = json.loads(line[len("LAUNCHER_RESULT "):])
    assert res["conv_class"] == "cim_quantization_b200.modules.lsq.Conv2dLSQCiM"
    assert len(res["epochs"]) == 1 and res["epochs"][0]["images"] == 192
    assert "after modules replacement" in r.stdout  # the reference's own surgery ran (examples/__init__.py:529)


@pytest.mark.gpu
@pytest.mark.skipif(not has_ref, reason="baseline/_ref not present (tools/make_baseline_ref.sh)")
def test_reference_checkpoint_round_trip(tmp_path):
    """SURVEY 8 f-4 (checkpoints): a `.pth.tar` written by the reference's save_checkpoint (examples/__init__.py:509-513)
    from a model built of our modules is loaded back by its `resume` logic (process_model, :540-548) into a fresh
    model: the validation pass main_lsq.py runs first reproduces the accuracy and loss of the end of the first run."""
    import json
    import re
    import subprocess

    def run(extra):
        r = subprocess.run([sys.executable, "-m", "cim_quantization_b200.launcher", "--impl", "ours", "--train-batches", "2",
                            "--val-batches", "1", "--batch-size", "64", "--workers", "0", "--work-dir", str(tmp_path)]
                           + extra, cwd=ROOT, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("LAUNCHER_RESULT ")][-1][16:])
        tests = re.findall(r"Test: \[0/1\].*?Loss (\S+) .*?Acc@1\s+(\S+)", r.stdout)
        return r.stdout, res, tests

    out1, res1, tests1 = run([])
    assert res1["checkpoints"], "the reference's save_checkpoint wrote nothing"
    ckpt = [c for c in res1["checkpoints"] if c.endswith("checkpoint.pth.tar")][-1]
    out2, res2, tests2 = run(["--set", f'resume="{ckpt}"'])
    assert "=> loading checkpoint" in out2
    # run 1: [validation before training, validation after the epoch]; run 2 starts from the saved weights
    assert len(tests1) == 2 and len(tests2) == 2
    assert tests2[0] == tests1[1], (tests1, tests2)


@pytest.mark.gpu
@pytest.mark.skipif(not has_ref, reason="baseline/_ref not present (tools/make_baseline_ref.sh)")
def test_reference_ddp_runs_on_our_kernels(tmp_path):
    """The reference's OWN multi-GPU path -- ``multi_gpu { multiprocessing_distributed: true }``: mp.spawn of
    main_worker per GPU, DistributedDataParallel, DistributedSampler (examples/__init__.py:80-104, 693-716) -- with
    ``models._modules`` replaced by this package in every spawned rank.  Needs two GPUs (``gpurun --gpus 2``)."""
    import json
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "cim_quantization_b200.launcher", "--impl", "ours", "--ddp",
                        "--train-batches", "4", "--val-batches", "1", "--batch-size", "128", "--workers", "0",
                        "--work-dir", str(tmp_path)], cwd=ROOT, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES="0,1"))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("LAUNCHER_RESULT ")][-1][16:])
    assert res.get("world_size") == 2 and len(res["ranks"]) == 2
    for rk in res["ranks"]:
        assert rk["conv_class"] == "cim_quantization_b200.modules.lsq.Conv2dLSQCiM"
        assert rk["epochs"][0]["model_class"] == "DistributedDataParallel"
        assert rk["epochs"][0]["images"] == 256  # DistributedSampler: half of the 512 training images per rank
