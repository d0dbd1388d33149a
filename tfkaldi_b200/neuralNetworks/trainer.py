"""Training environment with the reference's Trainer API (reference: neuralNetworks/trainer.py).

`Trainer.__init__` used to build a TF graph and every method was a `Session.run`; here the constructor
creates the CUDA engine and the methods are C-ABI calls:

    update_gradients_op.run(feed)                 -> tfk_accumulate        (trainer.py:165-169, 328)
    run([average_loss, apply_gradients_op]) + the three re-initialisers
                                                   -> tfk_apply             (trainer.py:174-184, 337-352)
    update_valid_loss.run(feed) / average_loss.eval() -> tfk_eval_accumulate / tfk_eval_finish
    halve_learningrate_op.run()                   -> tfk_halve_lr          (trainer.py:141-142)
    control_ops['add'] / ['init']                 -> tfk_set_active_layers / output-layer reset

Host batching keeps the reference's micro-batch structure (numutterances_per_minibatch utterances per
accumulate call) but packs the frames utterance-major into pinned memory instead of padding every
utterance to max_input_length (trainer.py:276-307) — the in-graph seq2nonseq stripped that padding
again anyway (classifiers/seq_convertors.py:12-39).
"""
from abc import ABCMeta, abstractmethod
import json
import os

import numpy as np
import torch

from .. import _lib as L
from ..engine import Engine
from ..processing.feeder import cmvn_coefficients


# reference TF variable name (below Classifier/layer<l>/) -> engine tensor stem
MODEL_NAMES = {"parameters/weights": "W", "parameters/biases": "b", "activation/batch_norm/beta": "beta",
               "activation/batch_norm/moving_mean": "moving_mean", "activation/batch_norm/moving_variance": "moving_var"}


def read_model_file(filename):
    """{variable name: array} from `<filename>.npz` (written here) or, when there is none, from the TensorFlow
    checkpoint the reference's savers leave at `filename` (V1 single file or V2 .index/.data; tf_checkpoint.py)."""
    path = filename if filename.endswith(".npz") else filename + ".npz"
    if os.path.exists(path):
        with np.load(path) as arrays:
            return {k: arrays[k] for k in arrays.files}
    from . import tf_checkpoint

    if tf_checkpoint.find(filename) is None:
        raise FileNotFoundError("neither %s nor a TensorFlow checkpoint at %s" % (path, filename))
    return tf_checkpoint.read(filename)


class _Op(object):
    """stand-in for a tf.Operation: something with .run()"""

    def __init__(self, fn):
        self._fn = fn

    def run(self):
        self._fn()


class _Stager(object):
    """Double-buffered pinned staging + side-stream H2D so packing batch k+1 overlaps computing batch k."""

    def __init__(self, device, input_dim, capacity):
        self.device, self.input_dim, self.capacity = device, input_dim, capacity
        self.host_x = [torch.empty((capacity, input_dim), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self.host_y = [torch.empty((capacity,), dtype=torch.int32, pin_memory=True) for _ in range(2)]
        self.dev_x = [torch.empty((capacity, input_dim), dtype=torch.float32, device=device) for _ in range(2)]
        self.dev_y = [torch.empty((capacity,), dtype=torch.int32, device=device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.turn = 0
        self.used = [False, False]

    def stage(self, mats, targets):
        """pack utterances (list of [T,I] / [T]) or one packed pair into the next buffer; returns device views"""
        k = self.turn
        self.turn ^= 1
        if self.used[k]:
            self.copied[k].synchronize()  # the previous H2D out of this pinned buffer has finished
        hx, hy = self.host_x[k].numpy(), self.host_y[k].numpy()
        if isinstance(mats, (list, tuple)):
            n = sum(m.shape[0] for m in mats)
            if n > self.capacity:
                raise ValueError("micro-batch of %d frames exceeds the staging capacity %d" % (n, self.capacity))
            np.concatenate(mats, axis=0, out=hx[:n])
            off = 0
            for t in targets:
                hy[off:off + t.shape[0]] = t  # uint32 -> int32 (placeholder dtype, trainer.py:52-55)
                off += t.shape[0]
            if off != n:
                raise ValueError("inputs hold %d frames but targets %d (CrossEnthropyTrainer needs equal lengths)" % (n, off))
        else:
            n = mats.shape[0]
            if n > self.capacity:
                raise ValueError("micro-batch of %d frames exceeds the staging capacity %d" % (n, self.capacity))
            hx[:n] = mats
            hy[:n] = targets
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            if self.used[k]:
                self.copy_stream.wait_event(self.consumed[k])  # compute no longer reads this device buffer
            self.dev_x[k][:n].copy_(self.host_x[k][:n], non_blocking=True)
            self.dev_y[k][:n].copy_(self.host_y[k][:n], non_blocking=True)
            self.copied[k].record(self.copy_stream)
        compute.wait_event(self.copied[k])
        self.used[k] = True
        return k, self.dev_x[k][:n], self.dev_y[k][:n]

    def stage_pinned(self, x, y, make_current_wait=True):
        """x [B, I] float32 / y [B] int32 torch tensors already in pinned host memory: H2D straight
        from the caller's buffers on the copy stream (no host-side repacking)"""
        k = self.turn
        self.turn ^= 1
        n = x.shape[0]
        if n > self.capacity:
            raise ValueError("micro-batch of %d frames exceeds the staging capacity %d" % (n, self.capacity))
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            if self.used[k]:
                self.copy_stream.wait_event(self.consumed[k])
            self.dev_x[k][:n].copy_(x, non_blocking=True)
            self.dev_y[k][:n].copy_(y, non_blocking=True)
            self.copied[k].record(self.copy_stream)
        if make_current_wait:
            compute.wait_event(self.copied[k])
        self.used[k] = True
        return k, self.dev_x[k][:n], self.dev_y[k][:n]

    def release(self, k):
        self.consumed[k].record(torch.cuda.current_stream(self.device))


class Trainer(object, metaclass=ABCMeta):
    """General training environment for a classifier (trainer.py:9-486)."""

    def __init__(self, classifier, input_dim, max_input_length, max_target_length, init_learning_rate,
                 learning_rate_decay, num_steps, numutterances_per_minibatch, *, precision="bf16", device=None,
                 seed=0, max_frames=None, distributed=False):
        self.classifier = classifier
        self.input_dim = input_dim
        self.numutterances_per_minibatch = numutterances_per_minibatch
        self.max_input_length = max_input_length
        self.max_target_length = max_target_length
        self.init_learning_rate = float(init_learning_rate)
        self.learning_rate_decay = float(learning_rate_decay)
        self.num_steps = int(num_steps)
        self.seed = seed
        spec = classifier.engine_spec(input_dim)
        self.loss_name = self.compute_loss()
        if max_frames is None:
            max_frames = int(numutterances_per_minibatch) * int(max_input_length)
        self.max_frames = int(max_frames)
        self.engine = Engine(spec["num_layers"], spec["input_dim"], spec["hidden_dim"], spec["output_dim"], self.max_frames,
                             nonlin=spec["nonlin"], batch_norm=spec["batch_norm"], keep_prob=spec["keep_prob"], l2_norm=spec["l2_norm"],
                             precision=precision, device=device, seed=seed)
        self.distributed = bool(distributed)
        self.rank, self.world = 0, 1
        if distributed:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
            self.engine.init_comm_from_torch()
        self._stager = None
        self.summarywriter = None
        # control operations for layer-by-layer initialisation (classifiers/dnn.py:81-120)
        if getattr(classifier, "layerwise_init", False):
            self.engine.set_active_layers(1)  # initialisedlayers = 0 -> logits from activations[0]
            self.control_ops = {"add": _Op(self._add_layer), "init": _Op(self._init_last_layer)}
        else:
            self.control_ops = None

    # ------------------------------------------------------------------ hooks
    @abstractmethod
    def compute_loss(self):
        """name of the loss kernel this trainer minimises (the reference builds the loss graph here,
        trainer.py:217-242)"""
        raise NotImplementedError("Abstract method")

    # ------------------------------------------------------------------ graph-level operations
    def initialize(self):
        """== init_op.run(): draw the initial parameters (classifiers/layer.py:39-48), zero everything else"""
        rng = np.random.default_rng(self.seed)
        self.engine.load_params(self.classifier.initial_parameters(self.input_dim, rng))
        zero = {"W": (L.T_ADAM_M_W, L.T_ADAM_V_W), "b": (L.T_ADAM_M_B, L.T_ADAM_V_B)}
        for l in range(self.engine.num_layers + 1):
            k, n = self.engine.layer_shape(l)
            for kind in zero["W"]:
                self.engine.set_tensor(kind, l, np.zeros((k, n), np.float32))
            for kind in zero["b"]:
                self.engine.set_tensor(kind, l, np.zeros(n, np.float32))
            if self.engine.batch_norm and l < self.engine.num_layers:
                for kind in (L.T_ADAM_M_BETA, L.T_ADAM_V_BETA):
                    self.engine.set_tensor(kind, l, np.zeros(n, np.float32))
        self.engine.set_scalar(L.S_GLOBAL_STEP, 0)
        self.engine.set_scalar(L.S_ADAM_STEP, 0)  # beta1_power / beta2_power back to their initial values (init_op)
        self.engine.set_scalar(L.S_LR_FACT, 1.0)

    def _add_layer(self):
        active = int(self.engine.get_scalar(L.S_ACTIVE_LAYERS))
        self.engine.set_active_layers(min(active + 1, self.engine.num_layers))

    def _init_last_layer(self):
        """tf.initialize_variables(<output layer scope>) (dnn.py:114-118): weights back to 0, biases to 0;
        the Adam slots are not part of that collection and keep their values."""
        l = self.engine.num_layers
        k, n = self.engine.layer_shape(l)
        self.engine.set_tensor(L.T_WEIGHTS, l, np.zeros((k, n), np.float32))
        self.engine.set_tensor(L.T_BIASES, l, np.zeros(n, np.float32))

    def start_visualization(self, logdir):
        """the reference opens a tf.train.SummaryWriter (trainer.py:249-258); we log the loss as JSON lines"""
        os.makedirs(logdir, exist_ok=True)
        self.summarywriter = open(os.path.join(logdir, "loss.jsonl"), "a")

    # ------------------------------------------------------------------ learning rate
    @property
    def global_step(self):
        return int(self.engine.get_scalar(L.S_GLOBAL_STEP))

    def learning_rate(self):
        """tf.train.exponential_decay(lr0, global_step, num_steps, decay), non-staircase (trainer.py:110-112);
        the learning_rate_fact factor is applied inside the engine"""
        return self.init_learning_rate * self.learning_rate_decay ** (float(self.global_step) / float(self.num_steps))

    # ------------------------------------------------------------------ batching
    def _stage(self, mats, targets):
        if self._stager is None:
            self._stager = _Stager(self.engine.device, self.input_dim, self.max_frames)
        return self._stager.stage(mats, targets)

    def _microbatches(self, inputs, targets):
        n = self.numutterances_per_minibatch
        if len(inputs) != len(targets):
            raise ValueError("inputs and targets hold a different number of utterances")
        if len(inputs) % n != 0:
            # the reference pads with (len % n) dummy utterances, which only yields whole micro-batches
            # when the remainder is 0 (trainer.py:280-294)
            raise ValueError("the number of utterances (%d) must be a multiple of numutterances_per_minibatch (%d)" % (len(inputs), n))
        for k in range(len(inputs) // n):
            yield inputs[k * n:(k + 1) * n], targets[k * n:(k + 1) * n]

    def update(self, inputs, targets):
        """update the model with a batch: list of [T_u, I] matrices + list of [T_u] target vectors;
        returns the mean loss per frame evaluated before the update (trainer.py:260-354)"""
        if len(inputs) == self.numutterances_per_minibatch:  # one micro-batch: fused accumulate+apply
            k, x, y = self._stage(inputs, targets)
            loss = self.engine.train_step(x, y, self.learning_rate_cached(), True)
            self._stager.release(k)
            return self._log(loss)
        for mats, tgts in self._microbatches(inputs, targets):
            k, x, y = self._stage(mats, tgts)
            self.engine.accumulate(x, y)
            self._stager.release(k)
        return self._apply()

    def prefetch(self, x, y):
        """start the host->device copy of the NEXT packed batch (pinned float32 / int32 tensors) on the
        copy stream so it overlaps the step in flight; the next update_packed(x, y) with the same tensors
        consumes it"""
        if self._stager is None:
            self._stager = _Stager(self.engine.device, self.input_dim, self.max_frames)
        k, dx, dy = self._stager.stage_pinned(x, y, make_current_wait=False)
        self._prefetched = (x.data_ptr(), y.data_ptr(), k, dx, dy)

    def update_packed(self, x, y, want_loss=True, prefetch=None):
        """fast path: one micro-batch already packed as x [B, I] float32 / y [B] int (pinned host tensors,
        numpy arrays or device tensors); same arithmetic as update()"""
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return self._log(self.engine.train_step(x, y, self.learning_rate_cached(), want_loss))
        else:
            if self._stager is None:
                self._stager = _Stager(self.engine.device, self.input_dim, self.max_frames)
            pre = getattr(self, "_prefetched", None)
            if pre is not None and isinstance(x, torch.Tensor) and pre[0] == x.data_ptr() and pre[1] == y.data_ptr():
                _, _, k, dx, dy = pre
                self._prefetched = None
                torch.cuda.current_stream(self.engine.device).wait_event(self._stager.copied[k])
            elif isinstance(x, torch.Tensor) and x.is_pinned() and y.is_pinned() and x.dtype == torch.float32 and y.dtype == torch.int32:
                k, dx, dy = self._stager.stage_pinned(x, y)
            else:
                if isinstance(x, torch.Tensor):
                    x, y = x.numpy(), y.numpy()
                k, dx, dy = self._stager.stage(x, y)
            # one micro-batch: fused step (per-layer Adam overlapped with the backward pass on one GPU).  The step is
            # launched first; the next batch's H2D is queued (copy stream) while it runs, then the host blocks on the loss
            lr = self.learning_rate_cached()
            if prefetch is None:
                loss = self.engine.train_step(dx, dy, lr, want_loss)
                self._stager.release(k)
                return self._log(loss)
            self.engine.train_step(dx, dy, lr, False)
            self._stager.release(k)
            self.prefetch(*prefetch)
            return self._log(self.engine.last_loss() if want_loss else None)

    def update_raw(self, raw_utts, cmvn_stats, targets, context_width):
        """update() from RAW features: list of un-normalised [T_u, D] matrices, their speakers' CMVN
        statistics ([2, D+1], processing/prepare_data.py:114-117) and target vectors.  CMVN and the
        +-context_width splice (FeatureReader.get_utt, feature_reader.py:42-60) run on the device, so
        the host uploads D instead of D*(2k+1) values per frame.  Utterances shorter than 2k+1 frames
        must already be filtered out, as the reference's dispenser does."""
        n = self.numutterances_per_minibatch
        if len(raw_utts) % n != 0:
            raise ValueError("the number of utterances (%d) must be a multiple of numutterances_per_minibatch (%d)" % (len(raw_utts), n))
        for k in range(len(raw_utts) // n):
            sl = slice(k * n, (k + 1) * n)
            mats, stats, tgts = raw_utts[sl], cmvn_stats[sl], targets[sl]
            lens = [m.shape[0] for m in mats]
            if min(lens) < 2 * context_width + 1:
                raise ValueError("utterance too short to splice")
            offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            cmvn = np.stack([cmvn_coefficients(st) for st in stats])  # apply_cmvn's formulas (feature_reader.py:109-115)
            labels = np.concatenate(tgts).astype(np.int32)
            self.engine.accumulate_raw(np.concatenate(mats, axis=0), offsets, cmvn, labels, mats[0].shape[1], context_width)
        return self._apply()

    def update_prefetched(self, feeder):
        """update() on the next batch of a RawBatchFeeder (processing/feeder.py): the batch was read and packed by the
        feeder's thread and its host->device copy queued while the previous step was computing; CMVN, splice and the
        whole step run on the device (tfk_train_step_raw, or accumulate_raw x k + apply for k micro-batches)."""
        batch = feeder.get_on_device(self.engine.device)
        lr, context = self.learning_rate_cached(), feeder.context_width
        parts = list(batch.microbatches())
        if len(parts) == 1:
            raw, labels, offsets, cmvn = parts[0]
            # launch the step, do the host work of the NEXT batch while it runs, only then block on the loss
            self.engine.train_step_raw(raw, offsets, cmvn, labels, batch.feat_dim, context, lr, False)
            feeder.consumed(batch)
            feeder.stage_next()
            return self._log(self.engine.last_loss())
        for raw, labels, offsets, cmvn in parts:
            self.engine.accumulate_raw(raw, offsets, cmvn, labels, batch.feat_dim, context)
        feeder.consumed(batch)
        self.engine.apply(self.learning_rate_cached(), False)
        feeder.stage_next()
        return self._log(self.engine.last_loss())

    def _apply(self, want_loss=True):
        return self._log(self.engine.apply(self.learning_rate_cached(), want_loss))

    def _log(self, loss):
        if self.summarywriter is not None and loss is not None:
            self.summarywriter.write(json.dumps({"step": self.global_step - 1, "loss": loss}) + "\n")
            self.summarywriter.flush()
        return loss

    def learning_rate_cached(self):
        if self.learning_rate_decay == 1.0:
            return self.init_learning_rate
        return self.learning_rate()

    def evaluate(self, inputs, targets):
        """mean loss per frame of the eval tower (moving-stat BN, no dropout) (trainer.py:356-441)"""
        if inputs is None or targets is None:
            return None
        for mats, tgts in self._microbatches(inputs, targets):
            k, x, y = self._stage(mats, tgts)
            self.engine.eval_accumulate(x, y)
            self._stager.release(k)
        return self.engine.eval_finish()

    def halve_learning_rate(self):
        self.engine.halve_lr()

    # ------------------------------------------------------------------ checkpoints
    # One .npz per reference checkpoint file, keyed by the reference's TF variable names.
    def _barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()

    def _writes_files(self):
        """data parallel: every rank runs the (collective) tensor reads, rank 0 alone touches the files"""
        return self.rank == 0

    def _sync_moving_stats(self, params):
        """Data parallel: a rank is one micro-batch of the step, and the reference updates the moving statistics once
        per micro-batch, one after the other (trainer.py:164-168).  The ranks run side by side, so each holds the EMA of
        its own micro-batches; the model that is written is their mean over ranks (all ranks call this)."""
        if self.world <= 1 or not self.engine.batch_norm:
            return params
        import torch.distributed as dist

        keys = sorted(k for k in params if k.startswith("moving_"))
        flat = torch.from_numpy(np.concatenate([params[k].ravel() for k in keys])).to(self.engine.device)
        dist.all_reduce(flat)
        flat = (flat / self.world).cpu().numpy()
        off = 0
        for k in keys:
            n = params[k].size
            params[k] = flat[off:off + n].reshape(params[k].shape).astype(np.float32)
            off += n
        return params

    def _model_arrays(self):
        out = {}
        for key, val in self._sync_moving_stats(self.engine.dump_params()).items():
            stem = key.rstrip("0123456789")
            layer = key[len(stem):]
            name = {v: k for k, v in MODEL_NAMES.items()}[stem]
            out["Classifier/layer%s/%s" % (layer, name)] = val
        if self.control_ops is not None:
            out["Classifier/initialisedlayers"] = np.array(int(self.engine.get_scalar(L.S_ACTIVE_LAYERS)) - 1, np.int32)
        return out

    def _load_model_arrays(self, arrays):
        params = {}
        for key in arrays:
            if key == "Classifier/initialisedlayers":
                self.engine.set_active_layers(int(arrays[key]) + 1)
                continue
            if not key.startswith("Classifier/"):
                continue
            _, layer, rest = key.split("/", 2)
            if rest not in MODEL_NAMES:
                raise ValueError("checkpoint variable %r is not part of the DNN classifier" % key)
            params[MODEL_NAMES[rest] + layer[len("layer"):]] = arrays[key]
        self.engine.load_params(params)

    @staticmethod
    def _path(filename):
        return filename if filename.endswith(".npz") else filename + ".npz"

    def save_model(self, filename):
        arrays = self._model_arrays()  # collective under data parallelism
        if self._writes_files():
            np.savez(self._path(filename), **arrays)
        self._barrier()  # the file exists when any rank returns (restore_* reads it on every rank)

    def restore_model(self, filename):
        """this engine's .npz, or a checkpoint written by the reference's tf.train.Saver (V1 or V2 format)"""
        self._load_model_arrays(read_model_file(filename))

    def export_tf_checkpoint(self, filename):
        """the model as a TensorFlow V2 checkpoint under the reference's variable names (tf.train.Saver().restore
        of the reference's graph reads it; trainer.py:456-463)"""
        from . import tf_checkpoint

        tf_checkpoint.write_v2(filename, self._model_arrays())

    def save_trainer(self, filename):
        """model + `train_variables` (global_step, learning_rate_fact) (trainer.py:465-475).  The Adam
        slots are written too (a superset: the reference never checkpoints them)."""
        self.save_model(filename)
        if self._writes_files():
            np.savez(self._path(filename + "_trainvars"), **{
                "train_variables/global_step": np.array(self.global_step, np.int32),
                "train_variables/learning_rate_fact": np.array(self.engine.get_scalar(L.S_LR_FACT), np.float32)})
        # Adam's step count (TF: beta1_power / beta2_power) belongs to the optimizer, not to train_variables
        slots = {"adam_step": np.array(int(self.engine.get_scalar(L.S_ADAM_STEP)), np.int64)}
        for l in range(self.engine.num_layers + 1):
            slots["W%d/Adam" % l] = self.engine.get_tensor(L.T_ADAM_M_W, l)
            slots["W%d/Adam_1" % l] = self.engine.get_tensor(L.T_ADAM_V_W, l)
            slots["b%d/Adam" % l] = self.engine.get_tensor(L.T_ADAM_M_B, l)
            slots["b%d/Adam_1" % l] = self.engine.get_tensor(L.T_ADAM_V_B, l)
            if self.engine.batch_norm and l < self.engine.num_layers:
                slots["beta%d/Adam" % l] = self.engine.get_tensor(L.T_ADAM_M_BETA, l)
                slots["beta%d/Adam_1" % l] = self.engine.get_tensor(L.T_ADAM_V_BETA, l)
        if self._writes_files():
            np.savez(self._path(filename + "_optimizer"), **slots)
        self._barrier()

    def restore_trainer(self, filename, restore_optimizer=False):
        """model + train_variables.  As in the reference the live Adam moments AND Adam's step count (beta powers)
        are left untouched (validation rollback keeps them, nnet.py:184-187; a resumed run starts from the freshly
        initialised optimizer, nnet.py:134-140) unless restore_optimizer=True."""
        self.restore_model(filename)
        tv = read_model_file(filename + "_trainvars")
        self.engine.set_scalar(L.S_GLOBAL_STEP, int(tv["train_variables/global_step"]))
        self.engine.set_scalar(L.S_LR_FACT, float(tv["train_variables/learning_rate_fact"]))
        if restore_optimizer:
            kinds = {"W": (L.T_ADAM_M_W, L.T_ADAM_V_W), "b": (L.T_ADAM_M_B, L.T_ADAM_V_B), "beta": (L.T_ADAM_M_BETA, L.T_ADAM_V_BETA)}
            with np.load(self._path(filename + "_optimizer")) as slots:
                for key in slots.files:
                    if key == "adam_step":
                        self.engine.set_scalar(L.S_ADAM_STEP, int(slots[key]))
                        continue
                    var, slot = key.split("/")
                    stem = var.rstrip("0123456789")
                    self.engine.set_tensor(kinds[stem][0 if slot == "Adam" else 1], int(var[len(stem):]), slots[key])


class CrossEnthropyTrainer(Trainer):
    """minimises the summed per-frame softmax cross-entropy against alignment targets (trainer.py:488-531)"""

    def compute_loss(self):
        return "softmax_cross_entropy_sum"
