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
