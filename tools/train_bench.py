"""ResNet-20 CiM training throughput (BASELINE.json configs 1/3/4/5) on synthetic CIFAR-shaped data.

    python tools/train_bench.py [--batch 256] [--steps 20] [--warmup 5] [--nbits 3] [--xbar 128] [--adcbits 1.5]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...   (data parallel)

One step = forward, cross-entropy, backward, flat NCCL all-reduce of all gradients (N > 1), SGD update.
Prints one JSON line (rank 0): images/s over all ranks, timed with CUDA events, max over ranks."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=256, help="images per GPU")
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--nbits", type=int, default=3)
    p.add_argument("--xbar", type=int, default=128)
    p.add_argument("--adcbits", type=float, default=1.5)
    p.add_argument("--width", type=int, default=1)
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--native-bn", action="store_true", help="ATen batch norm instead of cuDNN's (measured slower here)")
    p.add_argument("--torch-bn", action="store_true", help="torch / cuDNN batch norm instead of the fused kernels")
    a = p.parse_args()
    if a.native_bn:
        torch.backends.cudnn.enabled = False
    from cim_quantization_b200 import harness, _lib
    harness.FUSED_BN = not a.torch_bn
    from cim_quantization_b200.distributed import FlatGradAllReducer, broadcast_parameters
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    adc = int(a.adcbits) if a.adcbits == int(a.adcbits) else a.adcbits
    torch.manual_seed(0)
    model = harness.convert_to_cim(harness.resnet20(a.width), nbits_w=a.nbits, nbits_a=a.nbits, xbar=a.xbar,
                                   adcbits=adc).to(dev).train()
    torch.manual_seed(1 + rank)
    x = torch.randn(a.batch, 3, 32, 32, device=dev)
    y = torch.randint(0, 10, (a.batch,), device=dev)
    crit = torch.nn.CrossEntropyLoss()
    model(x)  # lazy initialisation of all step sizes on the first batch (lsq.py:532-563)
    broadcast_parameters(model, 0)
    opt = torch.optim.SGD(harness.sgd_param_groups(model), lr=0.01, momentum=0.9)
    reducer = FlatGradAllReducer(model.parameters())

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        if world > 1:
            reducer.all_reduce_()
        opt.step()
        return loss

    _lib.launch_counter = 0
    step()
    launches = _lib.launch_counter
    for _ in range(max(a.warmup, 3) - 1):
        step()
    torch.cuda.synchronize()
    runner, graphed = step, False
    if not a.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            g.replay()
            torch.cuda.synchronize()
            runner, graphed = g.replay, True
        except Exception as e:
            print(f"[train_bench] graph capture failed: {e!r}", file=sys.stderr)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        runner()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        nconv = sum(1 for m in model.modules() if m.__class__.__name__ == "Conv2dLSQCiM")
        print(json.dumps({"metric": "resnet20_cim_train_img_per_s", "value": a.batch * world * a.steps / (ms * 1e-3),
                          "unit": "img/s", "n_gpus": world, "steps": a.steps, "ms_per_step": ms / a.steps,
                          "scaling": "weak", "config": {"model": f"resnet20 width {a.width}", "cim_convs": nconv,
                                                        "w_a_bits": a.nbits, "xbar": a.xbar, "adcbits": adc,
                                                        "batch_per_gpu": a.batch, "global_batch": a.batch * world,
                                                        "cuda_graph": graphed, "optimizer": "SGD m0.9 wd1e-4"},
                          "gpu_launches_per_step": launches, "grad_allreduce_bytes": reducer.nbytes}))
    sys.stdout.flush()
    if world > 1:
        # the step graph holds NCCL work: tearing the communicator down afterwards can hang, and nothing is left
        # to do, so leave without the collective teardown
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
