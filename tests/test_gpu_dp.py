"""Data parallel on real GPUs (needs >= 2): K ranks, each feeding its own frames to tfk_accumulate and
calling tfk_apply, must equal ONE GPU accumulating the same shards as micro-batches
(trainer.py:310-332) — the semantics the reference has.  Three transports:
  fused         wgrad epilogue TMA-reduce-adds into the owner GPU's slice over NVLink peer memory, sharded Adam
                that stores the refreshed bf16 operands into every peer + flag publish/wait (default on a node)
  fused_nccl_ag the same with an NCCL all-gather of the operands
  sharded_nccl  NCCL reduce-scatter instead of the fused epilogue
  allreduce     NCCL all-reduce, replicated Adam"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(num_layers=2, input_dim=440, hidden_dim=256, output_dim=183, nonlin="linear")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _params():
    import math

    from oracle.dnn_oracle import OracleConfig, reference_init

    cfg = OracleConfig(**CFG)
    rng = np.random.default_rng(1)
    p = reference_init(cfg, rng)
    p["W2"] = (rng.standard_normal((256, 183)) / math.sqrt(256)).astype(np.float32)
    return p


def _shards(world):
    rng = np.random.default_rng(2)
    sizes = [200 + 56 * r for r in range(world)]  # unequal: the mean must use the GLOBAL frame count
    return [(rng.standard_normal((n, 440)).astype(np.float32), rng.integers(0, 183, n)) for n in sizes]


def _worker(rank, world, port, out_dir, mode):
    import torch
    import torch.distributed as dist

    from tfkaldi_b200.engine import Engine

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["TFK_DP_MODE"] = "" if mode == "fused" else mode
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = Engine(2, 440, 256, 183, 512, nonlin="linear", precision="bf16x3", device=rank)
    eng.load_params(_params())
    eng.init_comm_from_torch()
    x, y = _shards(world)[rank]
    losses = []
    for step in range(3):
        eng.accumulate(x + 0.1 * step, y)
        losses.append(eng.apply(1e-3))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), losses=np.array(losses), **eng.dump_params())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("mode", ["fused", "fused_nccl_ag", "sharded_nccl", "allreduce"])
def test_dp_equals_microbatch_accumulation(cuda_device, tmp_path, mode):
    import torch
    import torch.multiprocessing as mp

    from tfkaldi_b200.engine import Engine

    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = min(world, 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    single = Engine(2, 440, 256, 183, 512, nonlin="linear", precision="bf16x3", device=0)
    single.load_params(_params())
    losses = []
    for step in range(3):
        for x, y in _shards(world):
            single.accumulate(x + 0.1 * step, y)
        losses.append(single.apply(1e-3))
    want = single.dump_params()
    ranks = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    for r in ranks:
        assert np.allclose(r["losses"], losses, rtol=1e-5)
        for k, v in want.items():
            assert np.array_equal(r[k], ranks[0][k]), k  # replicas stay bit-identical
            assert np.abs(r[k] - v).max() <= 1e-3 * max(1.0, np.abs(v).max()), k
