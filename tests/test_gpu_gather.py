"""GPU tests (-m gpu) of the final gather over peer memory: two processes, one per rank, each
commits ITS slice of a batch with the fused kernel writing straight into rank 0's buffer (mapped through
lsr_peer_export / lsr_peer_open); rank 0 then holds the whole batch, bit for bit what one process computes and
what the oracle computes.  With two GPUs visible the ranks use cuda:0 / cuda:1 (stores cross NVLink); with one
GPU both ranks use cuda:0 -- the mapping, the slice arithmetic and the ABI are the same."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import Q0
from lambda_snark_r_b200 import sharding
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
SEED32 = bytes(range(32))
N, K, COUNT = 4096, 2, 12


def _worker(rank, world, port, ndev, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from lambda_snark_r_b200 import api, gather, sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % ndev
    torch.cuda.set_device(dev)
    api.set_device(dev)
    ctx = api.LweContext(api.Params(n=N, k=K, q=Q0, sigma=3.19), seed32=SEED32, validate=False)
    a, b = sh.shard_range(COUNT, rank, world)
    per = COUNT // world
    pg = gather.PeerGather(rank, world, per * ctx.words * 8, gather.torch_bcast())
    rng = np.random.Generator(np.random.PCG64(4242))
    msgs = rng.integers(0, 2**63, size=(COUNT, N), dtype=np.int64)          # same synthetic batch on all ranks
    dm = torch.from_numpy(msgs[a:b]).cuda()
    ds = torch.from_numpy(sh.global_seeds(0xC0FFEE, a, b).view(np.int64)).cuda()
    ctx.commit_batch_device(dm.data_ptr(), N, ds.data_ptr(), b - a, pg.slice_ptr, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()                                                          # every rank's stores have landed
    if rank == 0:
        full = gather.device_view(pg.base, COUNT * ctx.words).cpu().numpy().view(np.uint64).reshape(COUNT, ctx.words)
        np.save(Path(out_dir) / "gathered.npy", full)
    dist.barrier()
    pg.close()
    ctx.close()
    dist.destroy_process_group()


def test_two_ranks_commit_into_rank0_peer_memory(gpu, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() * 7) % 2000
    mp.spawn(_worker, args=(world, port, min(gpu, world), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    rng = np.random.Generator(np.random.PCG64(4242))
    msgs = rng.integers(0, 2**63, size=(COUNT, N), dtype=np.int64).view(np.uint64)
    want = O.OracleLwe(Q0, N, K, 3.19, SEED32).commit_batch(msgs, sharding.global_seeds(0xC0FFEE, 0, COUNT))
    assert np.array_equal(got, want)
