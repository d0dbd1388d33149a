"""Data-parallel host logic on CPU (world_size 2, gloo): a rank is one utterance micro-batch of
trainer.py:310-332, so  all-reduce(sum){grads, loss_sum, num_frames} -> /global frames -> clip -> Adam
on every rank must equal the single-process accumulation over both shards.  The CUDA engine does the
same reduction with NCCL inside tfk_apply; here the oracle stands in for the per-rank compute so the
sharding / reduction / bootstrap logic is exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.dnn_oracle import OracleConfig, OracleDNN, reference_init

CFG = dict(num_layers=2, input_dim=24, hidden_dim=32, output_dim=11)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(seed=0):
    cfg = OracleConfig(**CFG)
    rng = np.random.default_rng(seed)
    params = reference_init(cfg, rng)
    params["W2"] = (rng.standard_normal((32, 11)) / 6).astype(np.float32)
    return cfg, params


def _data():
    rng = np.random.default_rng(5)
    return rng.standard_normal((90, 24)).astype(np.float32), rng.integers(0, 11, 90)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # bootstrap exactly as Engine.init_comm_from_torch does: rank 0 creates 128 id bytes, all receive them
    payload = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    assert payload[0] == bytes(range(128))
    cfg, params = _make()
    orc = OracleDNN(cfg, params)
    x, y = _data()
    bounds = [0, 37, 90]  # unequal shards: the division must use the GLOBAL frame count
    for step in range(3):
        xs, ys = x[bounds[rank]:bounds[rank + 1]], y[bounds[rank]:bounds[rank + 1]]
        orc.accumulate(xs + step, ys)
        flat = torch.from_numpy(np.concatenate([orc.grads[k].ravel() for k in orc.trainable]))
        dist.all_reduce(flat)  # sum over ranks, like ncclAllReduce on the gradient arena
        scal = torch.tensor([orc.loss_sum, float(orc.num_frames)], dtype=torch.float64)
        dist.all_reduce(scal)
        off = 0
        for k in orc.trainable:
            n = orc.grads[k].size
            orc.grads[k][...] = flat[off:off + n].numpy().reshape(orc.grads[k].shape)
            off += n
        orc.loss_sum, orc.num_frames = float(scal[0]), int(scal[1])
        loss = orc.apply(1e-2)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss, **orc.p)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_dp_equals_microbatch_accumulation(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cfg, params = _make()
    ref = OracleDNN(cfg, params)
    x, y = _data()
    for step in range(3):
        ref.accumulate(x[:37] + step, y[:37])
        ref.accumulate(x[37:] + step, y[37:])
        loss = ref.apply(1e-2)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert abs(float(r0["loss"]) - loss) < 1e-6 and abs(float(r1["loss"]) - loss) < 1e-6
    for k in ref.p:
        assert np.array_equal(r0[k], r1[k]), k  # replicas stay bit-identical
        assert np.allclose(r0[k], ref.p[k], atol=2e-6), k
