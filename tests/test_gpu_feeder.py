"""The prefetching feeder end to end on the GPU: RawBatchFeeder -> Trainer.update_prefetched (tfk_train_step_raw /
tfk_accumulate_raw) must be the same arithmetic as the synchronous raw path (BatchDispenser.get_raw_batch ->
Trainer.update_raw) and agree with the reference-shaped host path (get_batch: host CMVN + splice -> Trainer.update),
including a validation rollback in the middle of the read-ahead."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def build(tmp_path, size, n, seed=7):
    from tfkaldi_b200 import synth
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer
    from tfkaldi_b200.processing import batchdispenser, feature_reader, target_coder

    info = synth.make_corpus(str(tmp_path / "corpus"), num_utts=40, min_len=30, max_len=90, feat_dim=40, num_speakers=4, num_pdfs=183, seed=2)
    fd = info["featdir"]

    def dispenser():
        reader = feature_reader.FeatureReader(fd + "/feats_shuffled.scp", fd + "/cmvn.scp", fd + "/utt2spk", 5, info["max_length"])
        return batchdispenser.AlignmentBatchDispenser(reader, target_coder.AlignmentCoder(lambda x, y: x, 183), size, info["alifile"])

    def trainer():
        dnn = DNN(183, 2, 256, act.Dropout(act.TfActivation(act.Batchnorm(None), act.relu), 0.8), False)
        tr = CrossEnthropyTrainer(dnn, 440, info["max_length"], info["max_length"], 1e-3, 1.0, 1000, n, precision="bf16", seed=seed)
        tr.initialize()
        # non-zero output layer so that every layer receives a gradient from the first step on
        from tfkaldi_b200 import _lib as L

        tr.engine.set_tensor(L.T_WEIGHTS, 2, (np.random.default_rng(seed).standard_normal((256, 183)) / 16).astype(np.float32))
        return tr

    return dispenser, trainer


@pytest.mark.parametrize("size,n", [(4, 4), (6, 2)])
def test_feeder_path_equals_synchronous_raw_path(cuda_device, tmp_path, size, n):
    from tfkaldi_b200.processing.feeder import RawBatchFeeder

    dispenser, trainer = build(tmp_path, size, n)
    a, b = trainer(), trainer()
    da, feeder = dispenser(), RawBatchFeeder(dispenser(), n)
    for step in range(2 * 40 // size + 2):  # two epochs and a bit: wrap-around, every slot reused several times
        a.engine.set_dropout_seed(500 + 10 * step)
        b.engine.set_dropout_seed(500 + 10 * step)
        mats, stats, targets = da.get_raw_batch()
        la = a.update_raw(mats, stats, targets, 5)
        lb = b.update_prefetched(feeder)
        assert la == lb, step
        if step == 5:  # validation got worse (nnet.py:177-186): two batches back on both sides
            for _ in range(2):
                da.return_batch()
                feeder.return_batch()
    pa, pb = a.engine.dump_params(), b.engine.dump_params()
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), k
    feeder.close()
    # same place in the scp (un-reading across the end of the list may leave the cursor at 0 instead of len: both
    # read utterance 0 next)
    ra, rb = feeder.dispenser.feature_reader.reader, da.feature_reader.reader
    assert ra.scp_position % len(ra.utt_ids) == rb.scp_position % len(rb.utt_ids)


def test_feeder_path_agrees_with_the_host_spliced_path(cuda_device, tmp_path):
    """device CMVN (x - mean) * (1/std) + splice vs the reference's host pipeline (x - mean) / std + splice: same losses
    to fp32 round-off in the fp32-equivalent mode (linear net, no dropout: nothing discontinuous)"""
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer
    from tfkaldi_b200.processing.feeder import RawBatchFeeder

    dispenser, _ = build(tmp_path, 4, 4)

    def trainer():
        tr = CrossEnthropyTrainer(DNN(183, 2, 256, act.TfActivation(None, act.linear), False), 440, 90, 90, 1e-3, 1.0, 1000, 4,
                                  precision="bf16x3", seed=3)
        tr.initialize()
        return tr

    a, b = trainer(), trainer()
    da, feeder = dispenser(), RawBatchFeeder(dispenser(), 4)
    for step in range(6):
        la = a.update(*da.get_batch())
        lb = b.update_prefetched(feeder)
        assert abs(la - lb) <= 1e-4 * max(1.0, abs(la)), (step, la, lb)
    feeder.close()


def test_update_packed_with_prefetch_equals_without(cuda_device, tmp_path):
    """update_packed(prefetch=next batch) launches the step, queues the next batch's copy and only then reads the loss
    (tfk_last_loss): identical losses and parameters to feeding the same pinned batches one at a time."""
    import torch

    _, trainer = build(tmp_path, 4, 4)
    a, b = trainer(), trainer()
    rng = np.random.default_rng(3)
    xs = [torch.from_numpy(rng.standard_normal((200 + 8 * i, 440)).astype(np.float32)).pin_memory() for i in range(5)]
    ys = [torch.from_numpy(rng.integers(0, 183, x.shape[0]).astype(np.int32)).pin_memory() for x in xs]
    b.prefetch(xs[0], ys[0])
    for i in range(5):
        a.engine.set_dropout_seed(900 + 10 * i)
        b.engine.set_dropout_seed(900 + 10 * i)
        la = a.update_packed(xs[i], ys[i])
        lb = b.update_packed(xs[i], ys[i], prefetch=(xs[i + 1], ys[i + 1]) if i < 4 else None)
        assert la == lb, i
    pa, pb = a.engine.dump_params(), b.engine.dump_params()
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), k
