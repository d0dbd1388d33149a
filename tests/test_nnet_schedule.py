"""Host logic of Nnet.train (reference: neuralNetworks/nnet.py:80-244) with a scripted trainer and dispenser: the
sequence of calls into the boundary — update / evaluate / save / restore / halve / control ops, and the dispenser's
cursor moves — must be the one the reference's loop makes.  Expected sequences are derived by hand from the reference
lines quoted in each test; no GPU and no arithmetic is involved (the fakes only record)."""
import configparser

import numpy as np
import pytest

import tfkaldi_b200.neuralNetworks.nnet as nnet_mod

CONF = """
[directories]
expdir = %(expdir)s
[nnet]
name = dnn
context_width = 5
num_hidden_units = 32
num_hidden_layers = %(layers)d
add_layer_period = %(add_layer_period)d
starting_step = %(starting_step)d
nonlin = relu
l2_norm = False
dropout = 1
batch_norm = False
num_epochs = %(epochs)d
initial_learning_rate = 0.001
learning_rate_decay = 1
batch_size = 4
numutterances_per_minibatch = %(nmb)s
valid_batches = %(valid_batches)d
valid_frequency = %(valid_frequency)d
valid_adapt = %(valid_adapt)s
valid_retries = %(valid_retries)d
check_freq = %(check_freq)d
visualise = False
"""


class Op(object):
    def __init__(self, log, name):
        self.log, self.name = log, name

    def run(self):
        self.log.append(self.name)


class ScriptedTrainer(object):
    """records every call; evaluate() plays back a list of validation losses"""
    log, losses, created = None, None, None

    def __init__(self, classifier, input_dim, max_input_length, max_target_length, init_learning_rate, learning_rate_decay,
                 num_steps, numutterances_per_minibatch, **kw):
        type(self).created = dict(input_dim=input_dim, max_input_length=max_input_length, max_target_length=max_target_length,
                                  lr=init_learning_rate, decay=learning_rate_decay, num_steps=num_steps, nmb=numutterances_per_minibatch)
        self.control_ops = {"add": Op(self.log, "add"), "init": Op(self.log, "init")}

    def initialize(self):
        self.log.append("initialize")

    def start_visualization(self, logdir):
        self.log.append("visualise")

    def update(self, inputs, targets):
        self.log.append("update:%s" % inputs[0])
        return 1.0

    def evaluate(self, inputs, targets):
        self.log.append("evaluate:%s" % ",".join(inputs))
        return self.losses.pop(0)

    def halve_learning_rate(self):
        self.log.append("halve")

    def save_trainer(self, f):
        self.log.append("save_trainer:" + f.rsplit("/", 1)[1])

    def restore_trainer(self, f):
        self.log.append("restore_trainer:" + f.rsplit("/", 1)[1])

    def save_model(self, f):
        self.log.append("save_model:" + f.rsplit("/", 1)[1])


class ScriptedDispenser(object):
    """batch k holds the single 'utterance' "b<k>"; the cursor moves like the reference's scp cursor"""

    def __init__(self, log, num_batches, size=4):
        self.log, self.pos, self.num_batches, self.size = log, 0, num_batches, size
        self.max_input_length, self.max_target_length = 777, 555

    def get_batch(self):
        self.pos += 1
        return ["b%d" % (self.pos - 1)], ["t%d" % (self.pos - 1)]

    def split(self):
        self.log.append("split@%d" % self.pos)

    def skip_batch(self):
        self.pos += 1
        self.log.append("skip")

    def return_batch(self):
        self.pos -= 1
        self.log.append("return")

    def compute_target_count(self):
        return np.array([1, 3, 0, 4])


def run(tmp_path, monkeypatch, losses, num_batches, **kw):
    opts = dict(expdir=str(tmp_path), layers=2, add_layer_period=0, starting_step=0, epochs=1, nmb="2", valid_batches=1,
                valid_frequency=2, valid_adapt="True", valid_retries=1, check_freq=3)
    opts.update(kw)
    conf = configparser.ConfigParser()
    conf.read_string(CONF % opts)
    log = []
    ScriptedTrainer.log, ScriptedTrainer.losses = log, list(losses)
    monkeypatch.setattr(nnet_mod, "CrossEnthropyTrainer", ScriptedTrainer)
    net = nnet_mod.Nnet(conf, 40, 4)
    net.train(ScriptedDispenser(log, num_batches))
    prior = np.load(str(tmp_path / "dnn" / "prior.npy"))
    assert prior.dtype == np.float32 and np.allclose(prior, [0.125, 0.375, 0, 0.5])  # nnet.py:241-244
    return log


def test_plain_schedule_validation_improves(tmp_path, monkeypatch):
    """nnet.py:88-99 (validation set = the first valid_batches batches, then split), :145-151 (initial validation +
    'validated' checkpoint), :154-166 (update, step += 1), :168-207 (every valid_frequency steps; better -> new
    validated checkpoint), :232-235 (check_freq checkpoints), :239 (final model)."""
    log = run(tmp_path, monkeypatch, losses=[5.0, 4.0, 3.0], num_batches=5)
    assert ScriptedTrainer.created == dict(input_dim=440, max_input_length=777, max_target_length=555, lr=0.001, decay=1.0, num_steps=5, nmb=2)
    assert log == ["split@1", "initialize", "evaluate:b0", "save_trainer:validated",
                   "update:b1", "update:b2", "evaluate:b0", "save_trainer:validated",
                   "update:b3", "save_trainer:step3",
                   "update:b4", "evaluate:b0", "save_trainer:validated",
                   "update:b5", "save_model:final"]


def test_rollback_halving_and_termination(tmp_path, monkeypatch):
    """nnet.py:177-200: a worse validation loss rewinds the dispenser by (step - validation_step) batches, restores the
    'validated' trainer, halves the learning rate and resumes from validation_step WITHOUT the check_freq checkpoint
    (`continue`); the restore + halve also happen on the attempt that terminates (num_retries == valid_retries)."""
    log = run(tmp_path, monkeypatch, losses=[5.0, 4.0, 4.5, 4.6], num_batches=8, check_freq=4)
    assert log == ["split@1", "initialize", "evaluate:b0", "save_trainer:validated",
                   "update:b1", "update:b2", "evaluate:b0", "save_trainer:validated",  # step 2: 4.0 < 5.0
                   "update:b3", "update:b4", "evaluate:b0",  # step 4: 4.5 > 4.0 (no step4 checkpoint: `continue`)
                   "return", "return", "restore_trainer:validated", "halve",
                   "update:b3", "update:b4", "evaluate:b0",  # the same two batches again
                   "return", "return", "restore_trainer:validated", "halve",  # second failure: retries exhausted
                   "save_model:final"]


def test_retry_counter_resets_after_an_improvement(tmp_path, monkeypatch):
    log = run(tmp_path, monkeypatch, losses=[5.0, 6.0, 4.0, 7.0, 3.0], num_batches=4, check_freq=100)
    assert log == ["split@1", "initialize", "evaluate:b0", "save_trainer:validated",
                   "update:b1", "update:b2", "evaluate:b0", "return", "return", "restore_trainer:validated", "halve",
                   "update:b1", "update:b2", "evaluate:b0", "save_trainer:validated",  # 4.0: num_retries back to 0
                   "update:b3", "update:b4", "evaluate:b0", "return", "return", "restore_trainer:validated", "halve",
                   "update:b3", "update:b4", "evaluate:b0", "save_trainer:validated",
                   "save_model:final"]


def test_valid_adapt_off_only_reports(tmp_path, monkeypatch):
    log = run(tmp_path, monkeypatch, losses=[5.0, 9.0], num_batches=3, valid_adapt="False", check_freq=100)
    assert log == ["split@1", "initialize", "evaluate:b0", "save_trainer:validated", "update:b1", "update:b2", "evaluate:b0",
                   "update:b3", "save_model:final"]


def test_layerwise_growth(tmp_path, monkeypatch):
    """nnet.py:210-229: every add_layer_period steps while step/period < num_hidden_layers: add, init, re-validate,
    new 'validated' checkpoint — after the regular validation of that step and before its check_freq checkpoint."""
    log = run(tmp_path, monkeypatch, losses=[5.0, 4.0, 3.5, 3.0], num_batches=6, layers=3, add_layer_period=2, valid_frequency=4, check_freq=2)
    assert log == ["split@1", "initialize", "evaluate:b0", "save_trainer:validated",
                   "update:b1", "update:b2", "add", "init", "evaluate:b0", "save_trainer:validated", "save_trainer:step2",  # 4.0
                   "update:b3", "update:b4", "evaluate:b0", "save_trainer:validated",  # step 4: 3.5 < 4.0
                   "add", "init", "evaluate:b0", "save_trainer:validated", "save_trainer:step4",  # 4/2 = 2 < 3 layers: grow, 3.0
                   "update:b5", "update:b6", "save_trainer:step6",  # 6/2 = 3: not < 3, no growth
                   "save_model:final"]


def test_resume_from_checkpoint_and_whole_batch_minibatches(tmp_path, monkeypatch):
    """nnet.py:101-108: starting_step is rounded DOWN to a multiple of check_freq, that many batches are skipped and
    training/step<k> is restored (:140-142); numutterances_per_minibatch = -1 means the dispenser's batch size (:110-114)."""
    log = run(tmp_path, monkeypatch, losses=[5.0, 4.0], num_batches=4, epochs=2, starting_step=5, check_freq=3, nmb="-1", valid_frequency=6)
    assert ScriptedTrainer.created["nmb"] == 4 and ScriptedTrainer.created["num_steps"] == 8
    assert log == ["split@1", "skip", "skip", "skip", "initialize", "restore_trainer:step3", "evaluate:b0", "save_trainer:validated",
                   "update:b4", "update:b5", "update:b6", "evaluate:b0", "save_trainer:validated", "save_trainer:step6",
                   "update:b7", "update:b8", "save_model:final"]


def test_unknown_nonlinearity_raises_like_the_reference(tmp_path):
    conf = configparser.ConfigParser()
    conf.read_string((CONF % dict(expdir=str(tmp_path), layers=2, add_layer_period=0, starting_step=0, epochs=1, nmb="2", valid_batches=1,
                                  valid_frequency=2, valid_adapt="True", valid_retries=1, check_freq=3)).replace("nonlin = relu", "nonlin = softsign"))
    with pytest.raises(Exception, match="unkown nonlinearity"):  # nnet.py:65 (sic)
        nnet_mod.Nnet(conf, 40, 4)
