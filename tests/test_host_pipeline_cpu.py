"""The host code above the C-ABI — Nnet.train / Nnet.decode / Trainer / Decoder / checkpoints / ark IO — end to end on
the CPU, with tests/oracle_engine.py standing in for the CUDA engine (the GPU twin of this file is
tests/test_gpu_nnet_e2e.py).  What is checked is ORCHESTRATION against the reference's (neuralNetworks/nnet.py:80-289,
trainer.py:260-486, decoder.py:49-81): which engine calls are made in which order, what lands in which file, and that
the decoded archive is what the saved weights produce — not kernel arithmetic."""
import configparser
import os

import numpy as np
import pytest

import tfkaldi_b200.neuralNetworks.decoder as decoder_mod
import tfkaldi_b200.neuralNetworks.trainer as trainer_mod
from oracle_engine import HostLane, HostStager, OracleEngine  # tests/oracle_engine.py (the tests directory is on sys.path)
from tfkaldi_b200 import _lib as L

NNET = """
[directories]
expdir = %(expdir)s
[nnet]
name = dnn
context_width = 5
num_hidden_units = 64
num_hidden_layers = 2
add_layer_period = %(add_layer_period)d
starting_step = 0
nonlin = relu
l2_norm = False
dropout = %(dropout)s
batch_norm = %(batch_norm)s
num_epochs = 2
initial_learning_rate = 0.001
learning_rate_decay = %(decay)s
batch_size = 8
numutterances_per_minibatch = 4
valid_batches = 1
valid_frequency = 3
valid_adapt = %(valid_adapt)s
valid_retries = 2
check_freq = 4
visualise = True
"""


@pytest.fixture
def host_only(monkeypatch):
    monkeypatch.setattr(trainer_mod, "Engine", OracleEngine)
    monkeypatch.setattr(trainer_mod, "_Stager", HostStager)
    monkeypatch.setattr(decoder_mod, "Engine", OracleEngine)
    monkeypatch.setattr(decoder_mod, "_Lane", HostLane)
    OracleEngine.calls = []
    return OracleEngine


def make_nnet(tmp_path, **kw):
    from tfkaldi_b200.neuralNetworks.nnet import Nnet

    opts = dict(expdir=str(tmp_path / "exp"), batch_norm="False", dropout="1", add_layer_period=0, decay="1", valid_adapt="True")
    opts.update(kw)
    os.makedirs(opts["expdir"], exist_ok=True)
    conf = configparser.ConfigParser()
    conf.read_string(NNET % opts)
    return Nnet(conf, 40, 60)


def corpora(tmp_path):
    from tfkaldi_b200 import synth
    from tfkaldi_b200.processing import batchdispenser, feature_reader, target_coder

    info = synth.make_corpus(str(tmp_path / "train"), num_utts=48, min_len=20, max_len=40, feat_dim=40, num_speakers=4, num_pdfs=60, seed=0)
    test = synth.make_corpus(str(tmp_path / "test"), num_utts=4, min_len=15, max_len=30, feat_dim=40, num_speakers=2, num_pdfs=60, seed=1, shuffle=False)
    fd = info["featdir"]
    reader = feature_reader.FeatureReader(fd + "/feats_shuffled.scp", fd + "/cmvn.scp", fd + "/utt2spk", 5, info["max_length"])
    dispenser = batchdispenser.AlignmentBatchDispenser(reader, target_coder.AlignmentCoder(lambda x, y: x, 60), 8, info["alifile"])
    treader = feature_reader.FeatureReader(test["featdir"] + "/feats.scp", test["featdir"] + "/cmvn.scp", test["featdir"] + "/utt2spk", 5, test["max_length"])
    return dispenser, treader, test


@pytest.mark.parametrize("batch_norm,dropout,add_layer_period", [("False", "1", 0), ("True", "0.8", 0), ("False", "1", 2)])
def test_train_then_decode_on_the_host(host_only, tmp_path, batch_norm, dropout, add_layer_period):
    from oracle.dnn_oracle import OracleConfig, OracleDNN
    from tfkaldi_b200.processing import ark

    dispenser, treader, test = corpora(tmp_path)
    # random labels: the validation loss never improves.  With valid_adapt the run rolls back twice and stops
    # (nnet.py:177-200); the layer-wise variant only reports the validation loss and trains all 12 steps.
    adapt = add_layer_period == 0
    nnet = make_nnet(tmp_path, batch_norm=batch_norm, dropout=dropout, add_layer_period=add_layer_period, valid_adapt=str(adapt))
    nnet.train(dispenser, prefetch=False)
    save = str(tmp_path / "exp" / "dnn")
    for f in ("final.npz", "prior.npy", "training/validated.npz", "training/validated_trainvars.npz", "training/validated_optimizer.npz",
              "logdir/loss.jsonl") + (() if adapt else ("training/step4.npz", "training/step8.npz", "training/step12.npz")):
        assert os.path.exists(os.path.join(save, f)), f
    assert host_only.calls.count("halve_lr") == (3 if adapt else 0)
    # every update = 8 utterances in micro-batches of 4: accumulate, accumulate, apply (trainer.py:310-346); every
    # validation = 8 utterances: eval_accumulate x 2, eval_finish (trainer.py:372-441)
    calls = [c for c in host_only.calls if c != "halve_lr"]
    i = 0
    while i < len(calls):
        if calls[i] == "accumulate":
            assert calls[i:i + 3] == ["accumulate", "accumulate", "apply"], calls[i:i + 4]
            i += 3
        else:
            assert calls[i:i + 3] == ["eval_accumulate", "eval_accumulate", "eval_finish"], calls[i:i + 4]
            i += 3
    import json

    steps = [json.loads(line)["step"] for line in open(save + "/logdir/loss.jsonl")]
    assert steps == ([0, 1, 2] * 3 if adapt else list(range(12)))  # global_step goes back with the restored trainer
    # decode: the archive must be log(softmax/prior) of the SAVED final weights
    decodedir = tmp_path / "decode"
    decodedir.mkdir()
    nnet.decode(treader, ark.ArkWriter(str(decodedir / "feats.scp"), str(decodedir / "likelihoods.ark")))
    final = np.load(save + "/final.npz")
    params, active = {}, 2
    for key in final.files:
        if key == "Classifier/initialisedlayers":
            active = int(final[key]) + 1
            continue
        _, layer, rest = key.split("/", 2)
        params[trainer_mod.MODEL_NAMES[rest] + layer[5:]] = final[key]
    if add_layer_period:
        assert active == 2  # grown to full depth by step 2 (1 layer at the start, +1 at step 2; 4/2 = 2 is not < 2)
    orc = OracleDNN(OracleConfig(2, 440, 64, 60, batch_norm=batch_norm == "True", keep_prob=float(dropout)), params)
    orc.active = active
    prior = np.load(save + "/prior.npy")
    out = ark.ArkReader(str(decodedir / "feats.scp"))
    assert out.utt_ids == test["utts"]
    from tfkaldi_b200.processing import feature_reader

    again = feature_reader.FeatureReader(test["featdir"] + "/feats.scp", test["featdir"] + "/cmvn.scp", test["featdir"] + "/utt2spk", 5, test["max_length"])
    for utt in test["utts"]:
        uid, x, _ = again.get_utt()
        got, want = out.read_utt(utt), orc.loglik(x, prior)
        assert uid == utt and got.dtype == np.float32 and got.shape == want.shape
        finite = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), finite) and np.abs(got[finite] - want[finite]).max() < 1e-5
    # the reference's loop shape (one utterance at a time through get_utt / write_next_utt) writes the same archive,
    # up to the fp32 round-off between host and "device" CMVN (the byte layout itself is checked in test_ark_streaming)
    plain = tmp_path / "decode_plain"
    plain.mkdir()
    treader.reader.scp_position = 0
    nnet.decode(treader, ark.ArkWriter(str(plain / "feats.scp"), str(plain / "likelihoods.ark")), streaming=False)
    assert open(plain / "feats.scp").read().replace(str(plain), "X") == open(decodedir / "feats.scp").read().replace(str(decodedir), "X")
    ref = ark.ArkReader(str(plain / "feats.scp"))
    for utt in test["utts"]:
        a, b = out.read_utt(utt), ref.read_utt(utt)
        ok = np.isfinite(b)
        assert a.shape == b.shape and np.abs(a[ok] - b[ok]).max() < 1e-4


def test_ark_streaming_writes_the_same_bytes(tmp_path):
    """ArkWriter.begin_utt / write_rows / finish_utt (the decoder's tile-by-tile output path) must produce exactly the
    archive and index write_next_utt does (processing/ark.py:190-211 of the reference), whatever the block order"""
    from tfkaldi_b200.processing import ark

    rng = np.random.default_rng(3)
    mats = {"uttA": rng.standard_normal((37, 11)).astype(np.float32), "b": rng.standard_normal((1, 11)).astype(np.float32),
            "utt-c": rng.standard_normal((4100, 11)).astype(np.float32)}
    w1 = ark.ArkWriter(str(tmp_path / "a.scp"), str(tmp_path / "a.ark"))
    for k, m in mats.items():
        w1.write_next_utt(k, m)
    w1.close()
    w2 = ark.ArkWriter(str(tmp_path / "b.scp"), str(tmp_path / "b.ark"))
    for i, (k, m) in enumerate(mats.items()):
        if i == 1:
            w2.write_next_utt(k, m)  # the two styles may be mixed in one archive
            continue
        e = w2.begin_utt(k, *m.shape)
        cuts = [0, 5, 6, 30, m.shape[0]] if m.shape[0] < 100 else [0, 1024, 3000, m.shape[0]]
        for lo, hi in reversed(list(zip(cuts[:-1], cuts[1:]))):  # out of order
            w2.write_rows(e, lo, m[lo:hi])
        w2.finish_utt(e)
    w2.close()
    assert open(tmp_path / "a.ark", "rb").read() == open(tmp_path / "b.ark", "rb").read()
    assert open(tmp_path / "a.scp").read().replace("a.ark", "b.ark") == open(tmp_path / "b.scp").read()


def test_trainer_host_logic(host_only, tmp_path):
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer

    rng = np.random.default_rng(0)
    dnn = DNN(30, 2, 32, act.TfActivation(act.Batchnorm(None), act.relu), True)  # layer-wise initialisation on
    tr = CrossEnthropyTrainer(dnn, 120, 50, 50, 0.01, 0.5, 10, 2, seed=1)
    tr.initialize()
    assert tr.control_ops is not None and tr.engine.get_scalar(L.S_ACTIVE_LAYERS) == 1  # initialisedlayers = 0 (dnn.py:85-92)
    x = [rng.standard_normal((t, 120)).astype(np.float32) for t in (20, 30, 25, 50)]
    y = [rng.integers(0, 30, t).astype(np.uint32) for t in (20, 30, 25, 50)]
    # learning rate: lr0 * decay^(global_step / num_steps), non-staircase (trainer.py:110-112)
    for step in range(3):
        assert tr.global_step == step and abs(tr.learning_rate() - 0.01 * 0.5 ** (step / 10.0)) < 1e-12
        host_only.calls.clear()
        tr.update(x, y)
        assert host_only.calls == ["accumulate", "accumulate", "apply"]  # 4 utterances, 2 per micro-batch
    host_only.calls.clear()
    tr.update(x[:2], y[:2])
    assert host_only.calls[0] == "train_step"  # exactly one micro-batch: the fused step
    with pytest.raises(ValueError, match="multiple of numutterances_per_minibatch"):
        tr.update(x[:3], y[:3])  # the reference's padding only works for whole micro-batches (App. A.11)
    with pytest.raises(ValueError):
        tr.update(x, y[:3])
    with pytest.raises(ValueError, match="exceeds the staging capacity"):
        tr.update([np.zeros((60, 120), np.float32), np.zeros((50, 120), np.float32)], [np.zeros(60, np.uint32), np.zeros(50, np.uint32)])
    assert tr.evaluate(None, None) is None
    # control ops (dnn.py:92, 114-118): add -> one more active layer (capped), init -> output layer back to zero
    tr.control_ops["add"].run()
    assert tr.engine.get_scalar(L.S_ACTIVE_LAYERS) == 2
    tr.control_ops["add"].run()
    assert tr.engine.get_scalar(L.S_ACTIVE_LAYERS) == 2
    assert np.abs(tr.engine.get_tensor(L.T_WEIGHTS, 2)).max() > 0
    m_before = tr.engine.get_tensor(L.T_ADAM_M_W, 2)
    tr.control_ops["init"].run()
    assert not tr.engine.get_tensor(L.T_WEIGHTS, 2).any() and not tr.engine.get_tensor(L.T_BIASES, 2).any()
    assert np.array_equal(tr.engine.get_tensor(L.T_ADAM_M_W, 2), m_before)  # Adam slots are not re-initialised
    # checkpoints: model + train variables + (superset) optimizer slots; rollback leaves the live moments alone
    tr.halve_learning_rate()
    path = str(tmp_path / "validated")
    tr.save_trainer(path)
    saved = tr.engine.dump_params()
    arrays = np.load(path + ".npz")
    assert int(arrays["Classifier/initialisedlayers"]) == 1  # 2 active layers
    assert sorted(k for k in arrays.files if "layer0" in k) == [
        "Classifier/layer0/activation/batch_norm/beta", "Classifier/layer0/activation/batch_norm/moving_mean",
        "Classifier/layer0/activation/batch_norm/moving_variance", "Classifier/layer0/parameters/biases", "Classifier/layer0/parameters/weights"]
    tv = np.load(path + "_trainvars.npz")
    assert int(tv["train_variables/global_step"]) == 4 and float(tv["train_variables/learning_rate_fact"]) == 0.5
    tr.update(x, y)
    tr.halve_learning_rate()
    m_live = tr.engine.get_tensor(L.T_ADAM_M_W, 0)
    tr.restore_trainer(path)
    assert tr.global_step == 4 and tr.engine.get_scalar(L.S_LR_FACT) == 0.5  # not 0.25: halving does not compound (App. B)
    assert all(np.array_equal(v, saved[k]) for k, v in tr.engine.dump_params().items())
    assert np.array_equal(tr.engine.get_tensor(L.T_ADAM_M_W, 0), m_live)  # the reference never checkpoints the moments
    tr.restore_trainer(path, restore_optimizer=True)
    assert not np.array_equal(tr.engine.get_tensor(L.T_ADAM_M_W, 0), m_live)
    # the reference's own checkpoint formats restore the same way (tf_checkpoint.py)
    tr.export_tf_checkpoint(str(tmp_path / "tfmodel"))
    tr.update(x, y)
    tr.restore_model(str(tmp_path / "tfmodel"))
    assert all(np.array_equal(v, saved[k]) for k, v in tr.engine.dump_params().items())


def test_decoder_host_logic(host_only, tmp_path):
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.decoder import Decoder
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer

    dnn = DNN(30, 1, 32, act.TfActivation(None, act.tanh), False)
    tr = CrossEnthropyTrainer(dnn, 120, 50, 50, 0.01, 1.0, 10, 1, seed=2)
    tr.initialize()
    rng = np.random.default_rng(1)
    tr.update([rng.standard_normal((40, 120)).astype(np.float32)], [rng.integers(0, 30, 40).astype(np.uint32)])
    tr.save_model(str(tmp_path / "final"))
    dec = Decoder(dnn, 120, 50)
    dec.restore(str(tmp_path / "final"))
    x = rng.standard_normal((17, 120)).astype(np.float32)
    post = dec(x)  # decoder.py:49-71: [T, O] softmax posteriors of the unpadded utterance
    assert post.shape == (17, 30) and post.dtype == np.float32 and np.abs(post.sum(1) - 1).max() < 1e-5
    assert np.allclose(post, tr.engine.orc.posteriors(x), atol=1e-7)


def test_adam_step_count_survives_restore_and_restarts_on_initialize(host_only, tmp_path):
    """tf.train.AdamOptimizer keeps beta1_power / beta2_power inside the optimizer (trainer.py:115); the
    `train_variables` saver holds global_step and learning_rate_fact only (trainer.py:204-205).  So
      * a validation rollback (restore_trainer, nnet.py:184-187) rewinds global_step but Adam keeps counting;
      * a resumed run (initialize() then restore_trainer(step N), nnet.py:134-140) restarts Adam at t = 1 with zero
        moments — NOT at t = N + 1, whose bias correction ~1 would make the first updates several times too large."""
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer

    def trainer():
        dnn = DNN(11, 2, 16, act.TfActivation(None, act.relu), False)
        tr = CrossEnthropyTrainer(dnn, 24, 50, 50, 1e-3, 1.0, 100, 1, seed=3)
        tr.initialize()
        return tr

    rng = np.random.default_rng(0)
    x = [rng.standard_normal((30, 24)).astype(np.float32)]
    y = [rng.integers(0, 11, 30).astype(np.uint32)]
    tr = trainer()
    for _ in range(5):
        tr.update(x, y)
    assert tr.global_step == 5 and tr.engine.get_scalar(L.S_ADAM_STEP) == 5
    tr.save_trainer(str(tmp_path / "step5"))
    for _ in range(3):
        tr.update(x, y)
    tr.restore_trainer(str(tmp_path / "step5"))  # rollback: global_step back to 5, Adam's t stays at 8
    assert tr.global_step == 5 and tr.engine.get_scalar(L.S_ADAM_STEP) == 8
    w_before = tr.engine.get_tensor(L.T_WEIGHTS, 2).copy()
    tr.update(x, y)
    assert tr.engine.get_scalar(L.S_ADAM_STEP) == 9

    fresh = trainer()  # resume in a new process: init_op, then restore_trainer
    fresh.restore_trainer(str(tmp_path / "step5"))
    assert fresh.global_step == 5 and fresh.engine.get_scalar(L.S_ADAM_STEP) == 0
    assert not fresh.engine.get_tensor(L.T_ADAM_M_W, 2).any()
    w0 = fresh.engine.get_tensor(L.T_WEIGHTS, 2).copy()
    fresh.update(x, y)
    step = np.abs(fresh.engine.get_tensor(L.T_WEIGHTS, 2) - w0).max()
    # Adam's first step from zero moments with t = 1 moves every weight by at most lr (sign-like); with t = 6 and
    # zero moments it would be lr * sqrt(1-b2^6)/(1-b1^6) * 0.1/sqrt(0.001) = 0.52 lr ... and at t >> 1 3.2 lr
    assert step <= 1.0001e-3, step
    full = trainer()  # restore_optimizer=True brings moments AND the step count back (superset of the reference)
    full.restore_trainer(str(tmp_path / "step5"), restore_optimizer=True)
    assert full.engine.get_scalar(L.S_ADAM_STEP) == 5 and full.engine.get_tensor(L.T_ADAM_V_W, 2).any()
    del w_before


@pytest.mark.parametrize("n", [8, 4])
def test_update_prefetched_defers_the_loss_read_behind_the_next_batchs_staging(host_only, tmp_path, n):
    """Trainer.update_prefetched: the step is launched WITHOUT a loss pointer, the slot is recycled and the next batch
    staged, and only then the loss is read (tfk_last_loss) — with the loss and the weights of the synchronous raw path
    (Trainer.update_raw), for one micro-batch per step (tfk_train_step_raw) and for two (accumulate_raw x 2 + apply)."""
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer
    from tfkaldi_b200.processing.feeder import RawBatchFeeder

    events = []

    class HostFeeder(RawBatchFeeder):  # the device half replaced: host tensors stand for the device views
        def get_on_device(self, device=None):
            batch = self._staged if self._staged is not None else self.get()
            self._staged = None
            return batch

        def stage_next(self):
            events.append("stage_next")
            if self._staged is None:
                self._staged = self.get()

        def consumed(self, batch):
            events.append("consumed")
            self._free.put(batch.slot)

    for name in ("train_step", "train_step_raw", "apply", "last_loss"):
        def wrap(self, *a, _f=getattr(OracleEngine, name), _n=name, **kw):
            events.append((_n, a[-1] if _n != "last_loss" else None))  # (call, want_loss)
            return _f(self, *a, **kw)
        setattr(host_only, "_orig_" + name, getattr(OracleEngine, name))
        setattr(host_only, name, wrap)
    try:
        da, _, _ = corpora(tmp_path)
        db, _, _ = corpora(tmp_path)

        def trainer():
            dnn = DNN(60, 2, 64, act.TfActivation(None, act.relu), False)
            tr = CrossEnthropyTrainer(dnn, 440, da.max_input_length, da.max_input_length, 1e-3, 1.0, 1000, n)
            tr.initialize()
            tr.engine.set_tensor(L.T_WEIGHTS, 2, (np.random.default_rng(5).standard_normal((64, 60)) / 8).astype(np.float32))
            return tr

        a, b = trainer(), trainer()
        feeder = HostFeeder(db, n)
        for step in range(3):
            la = a.update_raw(*da.get_raw_batch(), 5)
            del events[:]
            lb = b.update_prefetched(feeder)
            assert lb == la, step
            if n == 8:  # one micro-batch: the fused step (the stand-in's train_step_raw = train_step = accumulate + apply)
                want = [("train_step_raw", False), ("train_step", False), ("apply", False), "consumed"]
            else:  # two micro-batches: the batch's data has been consumed once both accumulates are queued
                want = ["consumed", ("apply", False)]
            assert events == want + ["stage_next", ("last_loss", None)], events
        feeder.close()
        for k, v in a.engine.dump_params().items():
            assert np.array_equal(v, b.engine.dump_params()[k]), k
        with pytest.raises(L.TfkError):
            b.engine.last_loss()  # read once
    finally:
        for name in ("train_step", "train_step_raw", "apply", "last_loss"):
            setattr(host_only, name, getattr(host_only, "_orig_" + name))
            delattr(host_only, "_orig_" + name)
