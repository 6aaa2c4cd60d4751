"""world_size-2 gloo tests of the multi-GPU plumbing (sharding, flat gradient all-reduce, max-over-ranks timing)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import ssmvs_b200
    from ssmvs_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    items = list(parallel.shard_items(5, r, w))
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    n = parallel.allreduce_gradients(lin.parameters())
    t = parallel.max_over_ranks(10.0 + rank)
    parallel.barrier()
    out[rank] = (items, n, lin.weight.grad.clone(), t)
    dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
    (i0, n0, g0, t0), (i1, n1, g1, t1) = out[0], out[1]
    assert i0 == [0, 1, 2] and i1 == [3, 4]                     # every item exactly once
    assert n0 == n1 == 15
    assert torch.allclose(g0, g1) and torch.allclose(g0, torch.full((3, 4), 3.0))  # mean of 2*1 and 2*2
    assert t0 == t1 == 11.0


def _flat_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import ssmvs_b200
    from ssmvs_b200 import parallel
    from ssmvs_b200.trainer import FlatGrads
    parallel.init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    flat = FlatGrads(net.parameters())
    opt = torch.optim.SGD(flat.params, lr=0.5)
    for step in range(2):                                   # gradients accumulate into the views; zero() clears them between steps
        flat.zero()
        net(torch.full((2, 4), float(rank + 1 + step))).sum().backward()
        assert all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in flat.params)       # autograd wrote INTO the flat buffer
        work = flat.allreduce()
        if work is not None:
            work.wait()
        opt.step()
    out[rank] = (flat.flat.clone(), [p.detach().clone() for p in net.parameters()])
    dist.destroy_process_group()


def test_flat_gradient_bucket_allreduce_keeps_replicas_identical():
    """trainer.FlatGrads: every .grad is a view of one flat fp32 buffer, ONE (async) all-reduce per optimiser step averages it
    over the ranks, and ranks that see different data stay bit-identical in their weights."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_flat_worker, args=(world, 29617, out), nprocs=world, join=True)
    (f0, p0), (f1, p1) = out[0], out[1]
    assert torch.equal(f0, f1) and f0.abs().sum() > 0
    assert all(torch.equal(a, b) for a, b in zip(p0, p1))


def _train_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    import ssmvs_b200
    from build_emu import build_emu
    from ssmvs_b200 import parallel, synth
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    from ssmvs_b200.trainer import TrainStep
    ssmvs_b200._lib.bind(build_emu())                       # CPU tensors: the host-emulation build of the SIMT kernels
    parallel.init_from_env("gloo")
    torch.manual_seed(0)                                    # the same initial weights on both ranks ...
    model = MVSNet(refine=False, train_dtype=torch.float32)
    step = TrainStep(model, UnSupLoss(), lr=1e-3)
    inp = synth.mvsnet_inputs(1, 4, 32, 64, 8, seed=10 + rank)                     # ... different data per rank (its shard of the batch)
    inp["imgs_aug"] = inp["imgs"] + 0.05 * torch.randn(inp["imgs"].shape, generator=torch.Generator().manual_seed(rank))
    before = [p.detach().clone() for p in model.parameters()]
    res = step(inp["imgs"], inp["imgs_aug"], inp["cams"], inp["proj_matrices"], inp["depth_values"])
    out[rank] = ([p.detach().clone() for p in model.parameters()], before, float(res["loss"]), step.grads.flat.clone())
    dist.destroy_process_group()


def test_two_rank_train_step_keeps_replicas_identical():
    """trainer.TrainStep under world_size 2 (gloo, host-emulation kernels): each rank runs train_sample + train_sample_aug on its
    own shard, the flat gradient bucket is all-reduced (averaged) before each of the two Adam steps, and the replicas end the batch
    with bit-identical weights although their data -- and losses -- differ (jdacs/train.py:65 does the same through DataParallel)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_train_worker, args=(world, 29619, out), nprocs=world, join=True)
    (p0, b0, l0, f0), (p1, b1, l1, f1) = out[0], out[1]
    assert all(torch.equal(a, b) for a, b in zip(b0, b1))                 # same start
    assert abs(l0 - l1) > 1e-6                                            # different shards
    assert all(torch.equal(a, b) for a, b in zip(p0, p1))                 # same end: the gradients were averaged
    assert torch.equal(f0, f1) and f0.abs().sum() > 0
    assert sum(int(not torch.equal(a, b)) for a, b in zip(p0, b0)) > 10   # and the step did train


def test_shard_items_covers_everything():
    from ssmvs_b200 import parallel
    for n in (1, 7, 8, 13):
        for w in (1, 2, 4, 8):
            got = [i for r in range(w) for i in parallel.shard_items(n, r, w)]
            assert got == list(range(n))


def test_numa_binding_is_best_effort():
    """Without NVML / a GPU the binding helper must change nothing and say so."""
    import os
    from ssmvs_b200 import parallel
    before = os.sched_getaffinity(0)
    ok = parallel.bind_to_gpu_numa(0)
    assert isinstance(ok, bool)
    if not ok:
        assert os.sched_getaffinity(0) == before


def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` launched the way the driver launches it for N > 1: rank 0 alone runs the CPU path and prints
    ONE JSON line with the contract's keys, the other rank exits 0 without work (no process group is created)."""
    import json
    import subprocess
    env = dict(os.environ, MVS_CPU_THREADS="4", OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and lines[0].startswith("{"), r.stdout[-2000:]    # native banners (NCCL's version line) go to stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "depth-samples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 4
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
