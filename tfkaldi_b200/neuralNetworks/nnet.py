"""Kaldi-style neural network façade with the reference's Nnet API (reference: neuralNetworks/nnet.py).

Same configuration keys ([nnet] section), same schedule (validation, learning-rate
halving with rollback, layer-wise growth, checkpoints, prior) and same files under
<expdir>/<name>/ as the reference; the TensorFlow sessions are replaced by the CUDA engine."""
import itertools
import os
import shutil

import numpy as np
import torch

from .classifiers import activation as act
from .classifiers.dnn import DNN
from .decoder import Decoder
from .trainer import CrossEnthropyTrainer


class Nnet(object):
    def __init__(self, conf, input_dim, num_labels, *, precision="bf16", device=None, distributed=False,
                 decode_precision="bf16x3"):
        """conf: ConfigParser with [nnet] and [directories] sections (nnet.py:17-78).

        precision: numeric mode of TRAINING ("bf16": single-pass tensor-core GEMMs, the timed mode; "bf16x3": fp32-
        equivalent).  decode_precision: numeric mode of decode(); the log-likelihoods Kaldi consumes carry the 1e-3
        contract against the reference's fp32 arithmetic (nnet.py:280-286, decoder.py:26-27), which only the
        fp32-equivalent mode meets, so that is the default."""
        self.decode_precision = decode_precision
        self.conf = dict(conf.items("nnet"))
        self.conf["savedir"] = conf.get("directories", "expdir") + "/" + self.conf["name"]
        os.makedirs(self.conf["savedir"] + "/training", exist_ok=True)
        self.precision, self.device, self.distributed = precision, device, distributed
        # spliced feature dimension
        self.input_dim = input_dim * (2 * int(self.conf["context_width"]) + 1)
        activation = act.Batchnorm(None) if self.conf["batch_norm"] == "True" else None
        nonlin = {"relu": act.relu, "sigmoid": act.sigmoid, "tanh": act.tanh, "linear": act.linear}.get(self.conf["nonlin"])
        if nonlin is None:
            raise Exception("unkown nonlinearity")
        activation = act.TfActivation(activation, nonlin)
        if self.conf["l2_norm"] == "True":
            activation = act.L2Norm(activation)
        if float(self.conf["dropout"]) < 1:
            activation = act.Dropout(activation, float(self.conf["dropout"]))
        self.dnn = DNN(num_labels, int(self.conf["num_hidden_layers"]), int(self.conf["num_hidden_units"]), activation,
                       int(self.conf["add_layer_period"]) > 0)

    def train(self, dispenser, prefetch=True):
        """train on the dispenser's data (nnet.py:80-244).

        prefetch=True (default): the dispenser is wrapped in a processing.feeder.RawBatchFeeder — a background thread
        reads and packs the RAW utterances of the next batches into pinned memory, CMVN + splicing run on the device
        (same utterances, same order, same cursor semantics for rollback/resume; values equal to the host pipeline to
        fp32 round-off).  prefetch=False keeps the reference's loop shape: get_batch() (host CMVN + splice) then
        update(), one after the other."""
        conf = self.conf
        if conf["numutterances_per_minibatch"] == "-1":
            numutterances_per_minibatch = dispenser.size
        else:
            numutterances_per_minibatch = int(conf["numutterances_per_minibatch"])
        if prefetch and hasattr(dispenser, "get_raw_batch") and not hasattr(dispenser, "get_on_device"):
            from ..processing.feeder import RawBatchFeeder

            dispenser = RawBatchFeeder(dispenser, numutterances_per_minibatch, device=self.device)
        val_data, val_labels = zip(*[dispenser.get_batch() for _ in range(int(conf["valid_batches"]))]) if int(conf["valid_batches"]) > 0 else ((), ())
        val_data = list(itertools.chain.from_iterable(val_data)) or None
        val_labels = list(itertools.chain.from_iterable(val_labels)) or None
        dispenser.split()
        num_steps = int(dispenser.num_batches * int(conf["num_epochs"]))
        step = int(conf["starting_step"]) - int(conf["starting_step"]) % int(conf["check_freq"])
        for _ in range(step):
            dispenser.skip_batch()
        trainer = CrossEnthropyTrainer(
            self.dnn, self.input_dim, dispenser.max_input_length, dispenser.max_target_length,
            float(conf["initial_learning_rate"]), float(conf["learning_rate_decay"]), num_steps,
            numutterances_per_minibatch, precision=self.precision, device=self.device, distributed=self.distributed)
        if conf["visualise"] == "True":
            if os.path.isdir(conf["savedir"] + "/logdir"):
                shutil.rmtree(conf["savedir"] + "/logdir")
            trainer.start_visualization(conf["savedir"] + "/logdir")
        trainer.initialize()
        if step > 0:
            trainer.restore_trainer(conf["savedir"] + "/training/step" + str(step))
        if val_data is not None:
            validation_loss = trainer.evaluate(val_data, val_labels)
            print("validation loss at step %d: %f" % (step, validation_loss))
            validation_step = step
            trainer.save_trainer(conf["savedir"] + "/training/validated")
            num_retries = 0
        prefetching = hasattr(dispenser, "get_on_device")  # a processing.feeder.RawBatchFeeder around the dispenser
        while step < num_steps:
            if prefetching:
                loss = trainer.update_prefetched(dispenser)
            else:
                batch_data, batch_labels = dispenser.get_batch()
                loss = trainer.update(batch_data, batch_labels)
            print("step %d/%d loss: %f" % (step, num_steps, loss))
            step += 1
            if step % int(conf["valid_frequency"]) == 0 and val_data is not None:
                current_loss = trainer.evaluate(val_data, val_labels)
                print("validation loss at step %d: %f" % (step, current_loss))
                if conf["valid_adapt"] == "True":
                    if current_loss > validation_loss:
                        # back to the validated model with half the learning rate (nnet.py:177-200)
                        for _ in range(step - validation_step):
                            dispenser.return_batch()
                        trainer.restore_trainer(conf["savedir"] + "/training/validated")
                        trainer.halve_learning_rate()
                        step = validation_step
                        if num_retries == int(conf["valid_retries"]):
                            print("the validation loss is worse, terminating training")
                            break
                        print("the validation loss is worse, returning to the previously validated model with halved learning rate")
                        num_retries += 1
                        continue
                    else:
                        validation_loss = current_loss
                        validation_step = step
                        num_retries = 0
                        trainer.save_trainer(conf["savedir"] + "/training/validated")
            period = int(conf["add_layer_period"])
            if period > 0 and step % period == 0 and step // period < int(conf["num_hidden_layers"]):
                print("adding layer, the model now holds %d/%d layers" % (step // period + 1, int(conf["num_hidden_layers"])))
                trainer.control_ops["add"].run()
                trainer.control_ops["init"].run()
                validation_loss = trainer.evaluate(val_data, val_labels)
                print("validation loss at step %d: %f" % (step, validation_loss))
                validation_step = step
                trainer.save_trainer(conf["savedir"] + "/training/validated")
                num_retries = 0
            if step % int(conf["check_freq"]) == 0:
                trainer.save_trainer(conf["savedir"] + "/training/step" + str(step))
        trainer.save_model(conf["savedir"] + "/final")
        self.trainer = trainer
        if prefetching:
            dispenser.close()  # stops the thread and un-reads what it had prefetched
        # state prior (nnet.py:241-244)
        prior = dispenser.compute_target_count().astype(np.float32)
        prior = prior / prior.sum()
        if getattr(trainer, "rank", 0) == 0:  # data parallel: one writer
            np.save(conf["savedir"] + "/prior.npy", prior)

    def decode(self, reader, writer, streaming=True):
        """pseudo log-likelihoods of every utterance of `reader` into `writer` (nnet.py:246-289).

        streaming=True (default): decoder.LoglikStreamer — raw features up (CMVN + splice on the device), the network and
        log(softmax / prior) on the device tile by tile, device->host copies and archive writes overlapped with the
        next tile.  Same archive bytes as streaming=False, which keeps the reference's loop: one utterance at a time,
        get_utt() (host CMVN + splice), evaluate, write_next_utt()."""
        from .decoder import LoglikStreamer

        decoder = Decoder(self.dnn, self.input_dim, reader.max_input_length, precision=self.decode_precision, device=self.device)
        prior = np.load(self.conf["savedir"] + "/prior.npy")
        decoder.restore(self.conf["savedir"] + "/final")
        if streaming and hasattr(reader, "get_raw_utt"):
            stream = LoglikStreamer(decoder, writer, prior.astype(np.float32))
            width = 1 + 2 * reader.context_width
            while True:
                utt_id, raw, stats, looped = reader.get_raw_utt()
                if looped:
                    break
                if raw.shape[0] < width:
                    raise ValueError("%s is too short to splice" % utt_id)  # the reference fails on utt_mat None too (nnet.py:277)
                stream.decode_raw(utt_id, raw, stats, reader.context_width)
            stream.close()
            writer.close()
            return
        prior_dev = torch.from_numpy(prior.astype(np.float32)).to(decoder.engine.device)
        while True:
            utt_id, utt_mat, looped = reader.get_utt()
            if looped:
                break
            # log(softmax/prior) on the device; the reference's flooring np.where is a no-op (nnet.py:283)
            loglik = decoder.loglik(utt_mat, prior_dev)
            writer.write_next_utt(utt_id, loglik.cpu().numpy())
        writer.close()
