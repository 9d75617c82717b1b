"""Multi-GPU paths on real devices (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):

* inference: `parallel.sharded_forward` — the global batch split over the ranks (whole T-windows), no data-path
  collective, gathered in rank order — equals the unsharded forward bit for bit;
* training: the ONE collective of the system — the bucketed ncclAllReduce(avg) of the gradients inside
  `w2l_wav2lip_train_step`, launched on a side stream as the backward completes each bucket: after a step on different
  data per rank, every rank holds the same averaged gradients (== the mean of the per-rank gradients computed without
  the collective) and the same parameters."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return dist


def _infer_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    try:
        from oracle import w2l_oracle as O
        from wav2lip_b200.models import Wav2Lip
        from wav2lip_b200.parallel import sharded_forward
        g = Wav2Lip()
        g.load_state_dict(O.make_state_dict("generator", 0), strict=True)
        g = g.cuda(rank).eval()
        mel, face = O.make_generator_inputs(6, seed=3, t=5)          # global batch B=6 windows of T=5
        mel, face = mel.cuda(rank), face.cuda(rank)
        with torch.no_grad():
            full = g(mel, face)
            sh = sharded_forward(lambda m, f: g(m, f), (mel, face))
        q.put((rank, bool(torch.equal(full, sh)), tuple(sh.shape)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _train_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    try:
        from oracle import w2l_oracle as O
        from wav2lip_b200.models import SyncNet_color, Wav2Lip
        from wav2lip_b200.training import Wav2LipTrainStep, init_data_parallel

        def make():
            g = Wav2Lip()
            g.load_state_dict(O.make_state_dict("generator", 0, init="default"), strict=True)
            e = SyncNet_color()
            e.load_state_dict(O.make_state_dict("syncnet", 1, init="default"), strict=True)
            return g.cuda(rank).train(), e.cuda(rank).train()

        gen = torch.Generator().manual_seed(50 + rank)                # different data on every rank
        B, T = 2, 5
        x = torch.rand((B, 6, T, 96, 96), generator=gen).cuda(rank)
        im = (torch.rand((B, T, 1, 80, 16), generator=gen) * 8 - 4).cuda(rank)
        mel = (torch.rand((B, 1, 80, 16), generator=gen) * 8 - 4).cuda(rank)
        gt = torch.rand((B, 3, T, 96, 96), generator=gen).cuda(rank)
        # (a) local gradients, no communicator
        g0, e0 = make()
        s0 = Wav2LipTrainStep(g0, e0, lr=1e-4, syncnet_wt=0.03)
        s0(x, im, mel, gt)
        local = s0.b.arena.clone()
        mean_of_locals = local.clone()
        dist.all_reduce(mean_of_locals)                               # torch's NCCL, as the independent reference
        mean_of_locals /= world
        # (b) the same step with the native bucketed all-reduce
        g1, e1 = make()
        s1 = Wav2LipTrainStep(g1, e1, lr=1e-4, syncnet_wt=0.03)
        assert init_data_parallel(s1) == world
        s1(x, im, mel, gt)
        torch.cuda.synchronize()
        got = s1.b.arena
        err = ((got - mean_of_locals).norm() / mean_of_locals.norm()).item()
        # parameters after the step are identical on every rank
        flat = torch.cat([p.detach().flatten() for p in g1.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        q.put((rank, err, bool(torch.equal(flat, ref)), float((local - got).norm() / got.norm())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _spawn(target, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_forward_equals_unsharded_on_gpus():
    for rank, same, shape in _spawn(_infer_worker, 2):
        assert same, rank
        assert shape == (6, 3, 5, 96, 96)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_training_step_all_reduce_averages_the_gradients():
    for rank, err, same_params, local_vs_avg in _spawn(_train_worker, 2):
        assert err <= 1e-5, (rank, err)               # fp32 sums of two ranks: NCCL's own arithmetic both ways
        assert same_params, rank
        assert local_vs_avg > 1e-2, rank              # the ranks really had different gradients before the collective
