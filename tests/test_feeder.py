"""Host half of the prefetching feeder (tfkaldi_b200/processing/feeder.py): what the background thread packs must be
exactly what the synchronous path (BatchDispenser.get_raw_batch -> Trainer.update_raw) feeds the device, in the same
utterance order, and the dispenser's cursor semantics (return_batch / skip_batch / split; reference
processing/batchdispenser.py:93-126, neuralNetworks/nnet.py:104-105,181-182) must survive the read-ahead."""
import time

import numpy as np
import pytest

from tfkaldi_b200 import synth
from tfkaldi_b200.processing import batchdispenser, feature_reader, target_coder
from tfkaldi_b200.processing.feeder import RawBatchFeeder, cmvn_coefficients
from tfkaldi_b200.processing.feature_reader import apply_cmvn


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    return synth.make_corpus(str(tmp_path_factory.mktemp("feed")), num_utts=37, min_len=11, max_len=60, feat_dim=13, num_speakers=3,
                             num_pdfs=50, seed=5)


def dispenser(info, size):
    fd = info["featdir"]
    reader = feature_reader.FeatureReader(fd + "/feats_shuffled.scp", fd + "/cmvn.scp", fd + "/utt2spk", 5, info["max_length"])
    return batchdispenser.AlignmentBatchDispenser(reader, target_coder.AlignmentCoder(lambda x, y: x, 50), size, info["alifile"])


def same(batch, raw_batch):
    mats, stats, targets = raw_batch
    assert batch.utts == len(mats) and batch.frames == sum(m.shape[0] for m in mats)
    assert np.array_equal(batch.raw.numpy(), np.concatenate(mats))
    assert batch.labels.dtype.is_floating_point is False and np.array_equal(batch.labels.numpy(), np.concatenate(targets).astype(np.int32))
    for i, st in enumerate(stats):
        assert np.array_equal(batch.cmvn.numpy()[i], cmvn_coefficients(st))
    return True


def test_cmvn_coefficients_reproduce_apply_cmvn(corpus):
    d = dispenser(corpus, 4)
    mats, stats, _ = d.get_raw_batch()
    for x, st in zip(mats, stats):
        c = cmvn_coefficients(st)
        assert c.dtype == np.float32 and c.shape == (2, 13)
        want = apply_cmvn(x, st)  # (x - mean) / sqrt(var), feature_reader.py:109-115
        assert np.abs((x - c[0]) * c[1] - want).max() < 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("size,n", [(4, 4), (6, 2), (5, 1)])
def test_batches_match_the_synchronous_path(corpus, size, n):
    feeder, twin = RawBatchFeeder(dispenser(corpus, size), n, depth=3), dispenser(corpus, size)
    assert feeder.feat_dim == 13 and feeder.context_width == 5
    for step in range(2 * 37 // size + 3):  # more than two epochs: the reader wraps around
        batch = feeder.get()
        want = twin.get_raw_batch()
        assert same(batch, want)
        parts = list(batch.microbatches())
        assert len(parts) == size // n
        row, utt = 0, 0
        for raw, labels, offsets, cmvn in parts:  # each micro-batch is self-contained: offsets restart at 0
            lens = [m.shape[0] for m in want[0][utt:utt + n]]
            assert offsets.tolist() == np.concatenate([[0], np.cumsum(lens)]).tolist()
            assert np.array_equal(raw.numpy(), np.concatenate(want[0][utt:utt + n])) and labels.shape[0] == sum(lens)
            assert cmvn.shape == (n, 2, 13)
            row, utt = row + sum(lens), utt + n
        assert row == batch.frames
        feeder.release(batch)
    feeder.close()
    # after close() the cursor sits right behind the last CONSUMED batch, whatever had been prefetched
    assert same_ids(feeder.dispenser, twin)


def same_ids(a, b):
    """exactly the same scp cursor (un-reading restores the recorded position, also across the end of the list)"""
    return a.feature_reader.reader.scp_position == b.feature_reader.reader.scp_position


def test_cursor_moves_survive_read_ahead(corpus):
    feeder, twin = RawBatchFeeder(dispenser(corpus, 4), depth=4), dispenser(corpus, 4)
    feeder.get_batch(), twin.get_batch()  # validation batch through the synchronous (spliced) path ...
    feeder.split(), twin.split()  # ... split off, as Nnet.train does (nnet.py:88-99)
    feeder.skip_batch(), twin.skip_batch()
    for _ in range(3):
        b = feeder.get()
        assert same(b, twin.get_raw_batch())
        feeder.release(b)
    time.sleep(0.2)  # let the thread run ahead as far as its slots allow
    assert feeder._ready.qsize() >= 2
    for _ in range(2):  # validation got worse: two batches back (nnet.py:181-182)
        feeder.return_batch(), twin.return_batch()
    for _ in range(4):
        b = feeder.get()
        assert same(b, twin.get_raw_batch())
        feeder.release(b)
    spliced, targets = feeder.get_batch()  # the synchronous path in the middle of prefetching
    want, want_t = twin.get_batch()
    assert all(np.array_equal(u, v) for u, v in zip(spliced, want)) and all(np.array_equal(u, v) for u, v in zip(targets, want_t))
    b = feeder.get()
    assert same(b, twin.get_raw_batch())
    feeder.release(b)
    feeder.close()
    assert same_ids(feeder.dispenser, twin)
    b = feeder.get()  # a closed feeder can be started again
    assert same(b, twin.get_raw_batch())
    feeder.release(b)
    feeder.close()
    assert feeder.num_batches == twin.num_batches and feeder.max_input_length == twin.max_input_length
    assert np.array_equal(feeder.compute_target_count(), twin.compute_target_count())


def test_errors_of_the_thread_reach_the_training_thread(corpus):
    d = dispenser(corpus, 4)
    feeder = RawBatchFeeder(d, capacity_frames=20)  # too small for any 4-utterance batch
    with pytest.raises(RuntimeError, match="feeder thread failed") as ei:
        feeder.get()
    assert "exceeds the feeder capacity" in str(ei.value.__cause__)
    with pytest.raises(ValueError, match="multiple of numutterances_per_minibatch"):
        RawBatchFeeder(d, utts_per_microbatch=3)


NNET_CONF = """
[directories]
expdir = %s
[nnet]
name = dnn
context_width = 5
num_hidden_units = 32
num_hidden_layers = 2
add_layer_period = 0
starting_step = 0
nonlin = relu
l2_norm = False
dropout = 1
batch_norm = False
num_epochs = 2
initial_learning_rate = 0.001
learning_rate_decay = 1
batch_size = 4
numutterances_per_minibatch = 2
valid_batches = 1
valid_frequency = 2
valid_adapt = True
valid_retries = 3
check_freq = 100
visualise = False
"""


class RecordingTrainer(object):
    """stands in for CrossEnthropyTrainer: records WHICH data every update consumed (frames, label checksum)"""
    log, losses = None, None

    def __init__(self, *a, **kw):
        self.control_ops = None

    def initialize(self):
        pass

    def update(self, inputs, targets):  # the reference's loop shape: spliced host batch
        self.log.append(("update", sum(m.shape[0] for m in inputs), int(sum(int(t.sum()) for t in targets))))
        return 1.0

    def update_prefetched(self, feeder):  # host half of Trainer.update_prefetched (no device copy here)
        batch = feeder.get()
        assert len(list(batch.microbatches())) == 2
        self.log.append(("update", batch.frames, int(batch.labels.numpy().sum())))
        feeder.release(batch)
        return 1.0

    def evaluate(self, inputs, targets):
        self.log.append(("evaluate", sum(m.shape[0] for m in inputs)))
        return self.losses.pop(0)

    def halve_learning_rate(self):
        self.log.append("halve")

    def save_trainer(self, f):
        self.log.append("save:" + f.rsplit("/", 1)[1])

    def restore_trainer(self, f):
        self.log.append("restore:" + f.rsplit("/", 1)[1])

    save_model = save_trainer


@pytest.mark.parametrize("losses", [[5, 4, 3, 2, 1, 0.5, 0.4, 0.3, 0.2, 0.1], [5, 4, 6, 3.5, 7, 8, 3, 2, 1, 0.5, 0.4, 0.3, 0.2]])
def test_nnet_train_consumes_the_same_data_with_and_without_prefetching(corpus, tmp_path, monkeypatch, losses):
    """Nnet.train(dispenser) wraps the dispenser in the feeder by default; with validation rollbacks in the middle of
    the read-ahead (nnet.py:177-186) every update must still see exactly the batch the synchronous loop sees, and
    the dispenser must end at the same place."""
    import configparser

    import tfkaldi_b200.neuralNetworks.nnet as nnet_mod

    monkeypatch.setattr(nnet_mod, "CrossEnthropyTrainer", RecordingTrainer)
    logs, ends = [], []
    for prefetch in (False, True):
        conf = configparser.ConfigParser()
        conf.read_string(NNET_CONF % str(tmp_path / ("p%d" % prefetch)))
        RecordingTrainer.log, RecordingTrainer.losses = [], list(losses)
        d = dispenser(corpus, 4)
        nnet_mod.Nnet(conf, 13, 50).train(d, prefetch=prefetch)
        logs.append(RecordingTrainer.log)
        ends.append(d.feature_reader.reader.scp_position)
    assert logs[0] == logs[1] and ends[0] == ends[1]
    assert sum(1 for e in logs[0] if e[0] == "update") >= 14  # 2 epochs x 7 batches (+ the repeated ones)
    if 6 in losses:
        assert "halve" in logs[0]
