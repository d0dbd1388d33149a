"""Nnet.train on two data-parallel ranks (gloo, CPU): every rank trains on its own shard of the corpus and validates on
its own validation utterances, so the rollback / learning-rate-halving / stop decisions of the reference's schedule
(neuralNetworks/nnet.py:168-207) must be taken on the loss SUMMED OVER RANKS, files must be written by one rank, and both
replicas must end with identical weights.  The per-rank compute is the oracle (tests/oracle_engine.py) with the two
reductions the CUDA engine performs with NCCL inside tfk_apply / tfk_eval_finish done over torch.distributed."""
import configparser
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

NNET = """
[directories]
expdir = %(expdir)s
[nnet]
name = dnn
context_width = 5
num_hidden_units = 48
num_hidden_layers = 2
add_layer_period = 0
starting_step = 0
nonlin = relu
l2_norm = False
dropout = 1
batch_norm = True
num_epochs = 2
initial_learning_rate = 0.001
learning_rate_decay = 1
batch_size = 4
numutterances_per_minibatch = 4
valid_batches = 1
valid_frequency = 2
valid_adapt = True
valid_retries = 1
check_freq = 3
visualise = False
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, root):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tfkaldi_b200.neuralNetworks.decoder as decoder_mod
    import tfkaldi_b200.neuralNetworks.trainer as trainer_mod
    from oracle_engine import HostLane, HostStager, OracleEngine
    from tfkaldi_b200 import synth
    from tfkaldi_b200.neuralNetworks.nnet import Nnet
    from tfkaldi_b200.processing import batchdispenser, feature_reader, target_coder

    decisions = []

    class DistOracleEngine(OracleEngine):
        """the oracle per rank + the reductions tfk_apply / tfk_eval_finish make under data parallelism"""

        def init_comm_from_torch(self):
            pass

        def apply(self, lr, want_loss=True):
            o = self.orc
            for k in o.trainable:
                t = torch.from_numpy(o.grads[k])
                dist.all_reduce(t)
            s = torch.tensor([o.loss_sum, float(o.num_frames)], dtype=torch.float64)
            dist.all_reduce(s)
            o.loss_sum, o.num_frames = float(s[0]), int(s[1])
            return super().apply(lr, want_loss)

        def eval_finish(self):
            o = self.orc
            s = torch.tensor([o.loss_sum, float(o.num_frames)], dtype=torch.float64)
            dist.all_reduce(s)
            o.loss_sum, o.num_frames = float(s[0]), int(s[1])
            out = super().eval_finish()
            decisions.append(round(out, 6))
            return out

    trainer_mod.Engine = DistOracleEngine
    trainer_mod._Stager = HostStager
    decoder_mod.Engine = DistOracleEngine
    decoder_mod._Lane = HostLane
    # every rank its own shard of the data (different seed => different utterances and validation set)
    info = synth.make_corpus(os.path.join(root, "train%d" % rank), num_utts=24, min_len=20, max_len=40, feat_dim=40, num_speakers=3,
                             num_pdfs=30, seed=10 + rank)
    fd = info["featdir"]
    reader = feature_reader.FeatureReader(fd + "/feats_shuffled.scp", fd + "/cmvn.scp", fd + "/utt2spk", 5, info["max_length"])
    dispenser = batchdispenser.AlignmentBatchDispenser(reader, target_coder.AlignmentCoder(lambda x, y: x, 30), 4, info["alifile"])
    conf = configparser.ConfigParser()
    conf.read_string(NNET % dict(expdir=os.path.join(root, "exp")))
    nnet = Nnet(conf, 40, 30, distributed=True)
    nnet.train(dispenser, prefetch=False)
    tr = nnet.trainer
    assert (tr.rank, tr.world) == (rank, world)
    np.savez(os.path.join(root, "result%d.npz" % rank), decisions=np.array(decisions), calls=np.array(OracleEngine.calls),
             **{k: v for k, v in tr.engine.dump_params().items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_takes_the_same_decisions_and_one_rank_writes(tmp_path):
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "exp"), exist_ok=True)
    mp.spawn(_worker, args=(2, _free_port(), root), nprocs=2, join=True)
    a, b = np.load(os.path.join(root, "result0.npz")), np.load(os.path.join(root, "result1.npz"))
    # identical validation losses on both ranks (they were summed over ranks) => identical schedules
    assert a["decisions"].shape == b["decisions"].shape and np.array_equal(a["decisions"], b["decisions"]) and len(a["decisions"]) >= 3
    assert list(a["calls"]) == list(b["calls"])
    assert "halve_lr" in list(a["calls"])  # random labels: the validation loss gets worse, the run rolls back at least once
    for k in a.files:
        if k in ("decisions", "calls"):
            continue
        if k.startswith("moving_"):
            continue  # batch-norm moving statistics are per rank (their mean over ranks is what save_model writes)
        assert np.array_equal(a[k], b[k]), k
    save = os.path.join(root, "exp", "dnn")
    for f in ("final.npz", "prior.npy", "training/validated.npz", "training/validated_trainvars.npz", "training/validated_optimizer.npz"):
        assert os.path.exists(os.path.join(save, f)), f
    final = np.load(os.path.join(save, "final.npz"))
    mm = final["Classifier/layer0/activation/batch_norm/moving_mean"]
    assert np.allclose(mm, 0.5 * (a["moving_mean0"] + b["moving_mean0"]), atol=1e-7)  # mean over the ranks
