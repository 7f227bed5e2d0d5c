"""HostBatchPipeline (harness.py): batches submitted from pinned host memory arrive on the GPU in order, the
slots are reused only after release(), and misuse raises."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_host_batch_pipeline_order_and_reuse():
    from cim_quantization_b200.harness import HostBatchPipeline
    dev = torch.device("cuda:0")
    pipe = HostBatchPipeline(dev, depth=2)
    host = [(torch.full((1 << 20,), float(i)).pin_memory(), torch.full((16,), float(-i)).pin_memory())
            for i in range(5)]
    got, submitted = [], 0
    with pytest.raises(RuntimeError):
        pipe.get()
    for n in range(5):
        while submitted < 5 and pipe.can_submit():
            pipe.submit(host[submitted])
            submitted += 1
        a, b = pipe.get()
        got.append((a.sum().item() / a.numel(), b[0].item()))  # consumed on the current stream
        pipe.release()
    assert got == [(float(i), float(-i)) for i in range(5)]
    pipe.submit(host[0])
    pipe.submit(host[1])
    assert not pipe.can_submit()
    with pytest.raises(RuntimeError):
        pipe.submit(host[2])
