"""End to end through the reference-facing API on a synthetic Kaldi corpus (config C1 shape):
AlignmentBatchDispenser -> Nnet.train (validation, LR halving + rollback, checkpoints, prior)
-> Nnet.decode -> ArkWriter; the log-likelihood archive is read back and compared with the oracle
run on the saved weights, and the archive bytes are checked against the reference's layout."""
import configparser
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NNET = """
[directories]
expdir = %(expdir)s
[nnet]
name = dnn
gmm_name = synthetic
context_width = 5
num_hidden_units = 256
num_hidden_layers = 2
add_layer_period = %(add_layer_period)d
starting_step = 0
monophone = True
nonlin = relu
l2_norm = False
dropout = %(dropout)s
batch_norm = %(batch_norm)s
num_epochs = 2
initial_learning_rate = 0.001
learning_rate_decay = 1
batch_size = 8
numutterances_per_minibatch = 4
valid_batches = 1
valid_frequency = 3
valid_adapt = True
valid_retries = 2
check_freq = 4
visualise = True
"""


def run_pipeline(tmp_path, batch_norm, dropout, add_layer_period):
    from tfkaldi_b200 import synth
    from tfkaldi_b200.neuralNetworks.nnet import Nnet
    from tfkaldi_b200.processing import ark, batchdispenser, feature_reader, target_coder

    info = synth.make_corpus(str(tmp_path / "train"), num_utts=48, min_len=60, max_len=120, feat_dim=40, num_speakers=4,
                             num_pdfs=183, seed=0)
    test = synth.make_corpus(str(tmp_path / "test"), num_utts=5, min_len=30, max_len=90, feat_dim=40, num_speakers=2,
                             num_pdfs=183, seed=1, shuffle=False)
    conf = configparser.ConfigParser()
    expdir = tmp_path / "exp"
    expdir.mkdir()
    conf.read_string(NNET % dict(expdir=str(expdir), batch_norm=batch_norm, dropout=dropout, add_layer_period=add_layer_period))
    nnet = Nnet(conf, 40, 183, precision="bf16x3")
    assert nnet.input_dim == 440
    featdir = info["featdir"]
    reader = feature_reader.FeatureReader(featdir + "/feats_shuffled.scp", featdir + "/cmvn.scp", featdir + "/utt2spk", 5, info["max_length"])
    coder = target_coder.AlignmentCoder(lambda x, y: x, 183)
    dispenser = batchdispenser.AlignmentBatchDispenser(reader, coder, 8, info["alifile"])
    nnet.train(dispenser)
    save = str(expdir / "dnn")
    for f in ("final.npz", "prior.npy", "training/validated.npz", "training/validated_trainvars.npz", "logdir/loss.jsonl"):
        assert os.path.exists(os.path.join(save, f)), f
    prior = np.load(save + "/prior.npy")
    assert prior.dtype == np.float32 and abs(prior.sum() - 1) < 1e-5 and prior.shape == (183,)
    # decode the test set
    treader = feature_reader.FeatureReader(test["featdir"] + "/feats.scp", test["featdir"] + "/cmvn.scp", test["featdir"] + "/utt2spk", 5, test["max_length"])
    decodedir = tmp_path / "decode"
    decodedir.mkdir()
    writer = ark.ArkWriter(str(decodedir / "feats.scp"), str(decodedir / "likelihoods.ark"))
    nnet.decode(treader, writer)
    return save, test, str(decodedir), prior


@pytest.mark.parametrize("batch_norm,dropout,add_layer_period", [("False", "1", 0), ("True", "0.8", 0), ("False", "1", 2)])
def test_train_decode_pipeline(cuda_device, tmp_path, batch_norm, dropout, add_layer_period):
    from oracle.dnn_oracle import OracleConfig, OracleDNN
    from tfkaldi_b200.processing import ark, feature_reader

    save, test, decodedir, prior = run_pipeline(tmp_path, batch_norm, dropout, add_layer_period)
    # the oracle on the SAVED weights must reproduce the decoded archive
    final = np.load(save + "/final.npz")
    names = {"parameters/weights": "W", "parameters/biases": "b", "activation/batch_norm/beta": "beta",
             "activation/batch_norm/moving_mean": "moving_mean", "activation/batch_norm/moving_variance": "moving_var"}
    params, active = {}, 2
    for key in final.files:
        if key == "Classifier/initialisedlayers":
            active = int(final[key]) + 1
            continue
        _, layer, rest = key.split("/", 2)
        params[names[rest] + layer[5:]] = final[key]
    cfg = OracleConfig(num_layers=2, input_dim=440, hidden_dim=256, output_dim=183, batch_norm=batch_norm == "True", keep_prob=float(dropout))
    orc = OracleDNN(cfg, params)
    orc.active = active
    out = ark.ArkReader(decodedir + "/feats.scp")
    treader = feature_reader.FeatureReader(test["featdir"] + "/feats.scp", test["featdir"] + "/cmvn.scp", test["featdir"] + "/utt2spk", 5, test["max_length"])
    assert out.utt_ids == test["utts"]
    raw = open(decodedir + "/likelihoods.ark", "rb").read()
    for i, utt in enumerate(test["utts"]):
        uid, x, _ = treader.get_utt()
        assert uid == utt
        got = out.read_utt(utt)
        want = orc.loglik(x, prior)
        assert got.dtype == np.float32 and got.shape == (test["lengths"][utt], 183)
        finite = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), finite)
        err = np.abs(got[finite] - want[finite]) / np.maximum(1, np.abs(want[finite]))
        assert err.max() < 1e-3, err.max()
        # byte layout (ark.py:204-210): key immediately followed by \0BFM, \4 rows, \4 cols, float32 payload
        pos = int(out.scp_data[i][1])
        assert raw[pos - len(utt):pos] == utt.encode() and raw[pos:pos + 5] == b"\0BFM "
        assert struct.unpack("<bibi", raw[pos + 5:pos + 15]) == (4, got.shape[0], 4, 183)


def test_tensorflow_checkpoints_through_the_trainer_and_decoder(cuda_device, tmp_path):
    """A model saved the way the reference saves it (tf.train.Saver: V1 single file or V2 index + data shard,
    trainer.py:448-486) restores through Trainer.restore_model / restore_trainer / Decoder.restore exactly like this
    engine's own .npz: bit-identical parameters, training variables, and posteriors."""
    from tfkaldi_b200.neuralNetworks import tf_checkpoint
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.decoder import Decoder
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer

    def make_dnn():
        return DNN(50, 2, 64, act.Dropout(act.TfActivation(act.Batchnorm(None), act.relu), 0.8), False)

    rng = np.random.default_rng(0)
    tr = CrossEnthropyTrainer(make_dnn(), 120, 64, 64, 1e-3, 1.0, 1000, 2, precision="bf16x3", seed=3)
    tr.initialize()
    x = [rng.standard_normal((t, 120)).astype(np.float32) for t in (40, 64)]
    y = [rng.integers(0, 50, t).astype(np.uint32) for t in (40, 64)]
    for _ in range(3):
        tr.update(x, y)
    tr.halve_learning_rate()
    want = tr.engine.dump_params()
    # (a) V2, written by export_tf_checkpoint; (b) V1, the r0.11 format, from the same arrays
    tr.export_tf_checkpoint(str(tmp_path / "v2model"))
    tf_checkpoint.write_v1(str(tmp_path / "v1model"), tr._model_arrays())
    trainvars = {"train_variables/global_step": np.array(3, np.int32), "train_variables/learning_rate_fact": np.array(0.5, np.float32),
                 "train_variables/num_frames": np.array(0, np.int32)}  # the reference's saver holds more than we need
    tf_checkpoint.write_v1(str(tmp_path / "v1model_trainvars"), trainvars)
    tr.save_model(str(tmp_path / "own"))
    probe = rng.standard_normal((33, 120)).astype(np.float32)
    outs = []
    for name in ("own", "v2model", "v1model"):
        dec = Decoder(make_dnn(), 120, 64, precision="bf16x3")
        dec.restore(str(tmp_path / name))
        got = dec.engine.dump_params()
        for k, v in want.items():
            assert np.array_equal(got[k], v), (name, k)
        outs.append(dec(probe))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    fresh = CrossEnthropyTrainer(make_dnn(), 120, 64, 64, 1e-3, 1.0, 1000, 2, precision="bf16x3", seed=99)
    fresh.initialize()
    fresh.restore_trainer(str(tmp_path / "v1model"))
    assert fresh.global_step == 3
    from tfkaldi_b200 import _lib as L

    assert fresh.engine.get_scalar(L.S_LR_FACT) == 0.5
    got = fresh.engine.dump_params()
    assert all(np.array_equal(got[k], v) for k, v in want.items())
    with pytest.raises(FileNotFoundError):
        fresh.restore_model(str(tmp_path / "nothing_here"))


def test_streaming_decode_tiles_and_bytes(cuda_device, tmp_path):
    """decoder.LoglikStreamer (what Nnet.decode runs): utterances longer than the tile, so every utterance crosses
    several device tiles, pinned slots are recycled under back-pressure (2 slots, 3 writer threads) and utterances
    overlap.  The archive must hold exactly what one whole-utterance tfk_forward_loglik_raw call returns, in the
    reference's byte layout (ark.py:204-210), indexed in utterance order."""
    import torch

    from oracle.dnn_oracle import OracleConfig, reference_init
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.decoder import Decoder, LoglikStreamer
    from tfkaldi_b200.processing import ark
    from tfkaldi_b200.processing.feeder import cmvn_coefficients

    rng = np.random.default_rng(11)
    dnn = DNN(183, 2, 256, act.TfActivation(None, act.relu), False)
    dec = Decoder(dnn, 440, 4000, max_frames=512)
    assert dec.engine.precision == "bf16x3"  # log-likelihoods default to the fp32-equivalent mode
    params = reference_init(OracleConfig(2, 440, 256, 183), rng)
    params["W2"] = (rng.standard_normal((256, 183)) / 16).astype(np.float32)
    dec.engine.load_params(params)
    prior = (rng.random(183) + 0.1).astype(np.float32)
    prior /= prior.sum()
    utts = {"utt%d" % i: (3.0 + rng.standard_normal((n, 40))).astype(np.float32) for i, n in enumerate([1500, 11, 700, 512, 513])}
    stats = np.zeros((2, 41), np.float32)
    stats[0, :-1], stats[0, -1], stats[1, :-1] = 3.0 * 500, 500, (9.0 + 1.3) * 500
    writer = ark.ArkWriter(str(tmp_path / "feats.scp"), str(tmp_path / "ll.ark"))
    stream = LoglikStreamer(dec, writer, prior, tile=512, slots=2, io_threads=3)
    for k, m in utts.items():
        stream.decode_raw(k, m, stats, 5)
    stream.close()
    writer.close()
    out = ark.ArkReader(str(tmp_path / "feats.scp"))
    assert out.utt_ids == list(utts)
    raw = open(tmp_path / "ll.ark", "rb").read()
    coef = cmvn_coefficients(stats)[None]
    for i, (k, m) in enumerate(utts.items()):
        want = dec.engine.loglik_raw(m, np.array([0, m.shape[0]], np.int32), coef, 40, 5, prior).cpu().numpy()
        got = out.read_utt(k)
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-5, k
        pos = int(out.scp_data[i][1])
        assert raw[pos - len(k):pos] == k.encode() and raw[pos:pos + 5] == b"\0BFM "
        assert struct.unpack("<bibi", raw[pos + 5:pos + 15]) == (4, m.shape[0], 4, 183)
    assert len(raw) == sum(len(k) + 15 + m.shape[0] * 183 * 4 for k, m in utts.items())
    torch.cuda.synchronize()
