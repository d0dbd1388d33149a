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
