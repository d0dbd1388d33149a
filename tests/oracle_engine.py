"""TEST INFRASTRUCTURE: a stand-in for tfkaldi_b200.engine.Engine whose arithmetic is the CPU oracle.

It exists so that the HOST code above the C-ABI (Trainer, Nnet, Decoder, checkpoints, staging, ark IO) can be exercised
by the CPU test suite (`-m "not gpu"`), where no CUDA device exists.  tests/test_host_pipeline_cpu.py monkeypatches it in;
the product never imports this module (tests/test_cabi_surface.py::test_product_never_imports_the_oracle scans the
package) and has no CPU path of its own."""
import numpy as np
import torch

from oracle.dnn_oracle import OracleConfig, OracleDNN, reference_init
from tfkaldi_b200 import _lib as L
from tfkaldi_b200.processing.feature_reader import splice


def _np(x, dtype=np.float32):
    if isinstance(x, torch.Tensor):
        x = x.cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


class OracleEngine(object):
    calls = []  # (method name) log shared by every instance: the tests read the call ORDER from it

    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, max_frames, *, nonlin="relu", batch_norm=False,
                 keep_prob=1.0, precision="bf16", device=None, seed=0, l2_norm=False):
        if nonlin not in ("relu", "linear", "sigmoid", "tanh"):
            raise Exception("unkown nonlinearity")
        self.cfg = OracleConfig(num_layers, input_dim, hidden_dim, output_dim, nonlin=nonlin, batch_norm=batch_norm,
                                keep_prob=keep_prob, l2_norm=l2_norm)
        self.orc = OracleDNN(self.cfg, reference_init(self.cfg, np.random.default_rng(0)))
        self.num_layers, self.input_dim, self.hidden_dim, self.output_dim = num_layers, input_dim, hidden_dim, output_dim
        self.max_frames, self.batch_norm, self.precision = int(max_frames), bool(batch_norm), precision
        self.device = torch.device("cpu")
        self.drop_seed = int(seed)

    # ---------------------------------------------------------------- tensors / scalars
    def layer_shape(self, layer):
        return (self.input_dim if layer == 0 else self.hidden_dim), (self.hidden_dim if layer < self.num_layers else self.output_dim)

    def _slot(self, kind, layer):
        o = self.orc
        table = {L.T_WEIGHTS: (o.p, "W"), L.T_BIASES: (o.p, "b"), L.T_BN_BETA: (o.p, "beta"), L.T_BN_MOVING_MEAN: (o.p, "moving_mean"),
                 L.T_BN_MOVING_VAR: (o.p, "moving_var"), L.T_ADAM_M_W: (o.m, "W"), L.T_ADAM_V_W: (o.v, "W"), L.T_ADAM_M_B: (o.m, "b"),
                 L.T_ADAM_V_B: (o.v, "b"), L.T_ADAM_M_BETA: (o.m, "beta"), L.T_ADAM_V_BETA: (o.v, "beta"), L.T_GRAD_W: (o.grads, "W"),
                 L.T_GRAD_B: (o.grads, "b"), L.T_GRAD_BETA: (o.grads, "beta")}
        store, stem = table[kind]
        return store, "%s%d" % (stem, layer)

    def set_tensor(self, kind, layer, array):
        store, key = self._slot(kind, layer)
        if key not in store or store[key].size != np.asarray(array).size:
            raise ValueError("tensor kind %d layer %d" % (kind, layer))
        store[key][...] = np.asarray(array, np.float32).reshape(store[key].shape)

    def get_tensor(self, kind, layer):
        store, key = self._slot(kind, layer)
        return store[key].copy()

    def set_scalar(self, kind, value):
        if kind == L.S_GLOBAL_STEP:
            self.orc.global_step = int(value)
        elif kind == L.S_ADAM_STEP:
            self.orc.adam_step = int(value)
        elif kind == L.S_LR_FACT:
            self.orc.lr_fact = float(value)
        else:
            raise ValueError(kind)

    def get_scalar(self, kind):
        o = self.orc
        return float({L.S_GLOBAL_STEP: o.global_step, L.S_LR_FACT: o.lr_fact, L.S_ACTIVE_LAYERS: o.active, L.S_LOSS_SUM: o.loss_sum,
                      L.S_NUM_FRAMES: o.num_frames, L.S_ADAM_STEP: o.adam_step}[kind])

    def load_params(self, params):
        for key, val in params.items():
            self.orc.p[key][...] = np.asarray(val, np.float32).reshape(self.orc.p[key].shape)

    def dump_params(self):
        return {k: v.copy() for k, v in self.orc.p.items()}

    # ---------------------------------------------------------------- steps
    def _check(self, n):
        if n <= 0 or n > self.max_frames:
            raise L.TfkError(L.TFK_ESHAPE, "B=%d exceeds max_frames=%d" % (n, self.max_frames))

    def accumulate(self, x, labels):
        self.calls.append("accumulate")
        x = _np(x)
        self._check(x.shape[0])
        self.orc.accumulate(x, _np(labels, np.int64), dropout_seed=self.drop_seed)
        self.drop_seed += self.num_layers + 1

    def apply(self, lr, want_loss=True):
        self.calls.append("apply")
        loss = self.orc.apply(lr)
        self._pending_loss = None if want_loss else float(loss)
        return float(loss) if want_loss else None

    def last_loss(self):
        if getattr(self, "_pending_loss", None) is None:
            raise L.TfkError(L.TFK_EINVAL, "tfk_last_loss: no step with an unread loss is outstanding")
        loss, self._pending_loss = self._pending_loss, None
        return loss

    def train_step(self, x, labels, lr, want_loss=True):
        self.calls.append("train_step")
        self.accumulate(x, labels)
        return self.apply(lr, want_loss)

    def _spliced(self, raw, utt_offsets, cmvn, feat_dim, context):
        raw, off, cmvn = _np(raw), _np(utt_offsets, np.int64), _np(cmvn)
        return np.concatenate([splice((raw[off[u]:off[u + 1]] - cmvn[u, 0]) * cmvn[u, 1], context) for u in range(len(off) - 1)])

    def accumulate_raw(self, raw, utt_offsets, cmvn, labels, feat_dim, context):
        self.accumulate(self._spliced(raw, utt_offsets, cmvn, feat_dim, context), labels)

    def train_step_raw(self, raw, utt_offsets, cmvn, labels, feat_dim, context, lr, want_loss=True):
        return self.train_step(self._spliced(raw, utt_offsets, cmvn, feat_dim, context), labels, lr, want_loss)

    def eval_accumulate(self, x, labels):
        self.calls.append("eval_accumulate")
        x = _np(x)
        self._check(x.shape[0])
        self.orc.eval_accumulate(x, _np(labels, np.int64))

    def eval_finish(self):
        self.calls.append("eval_finish")
        return float(self.orc.eval_finish())

    def posteriors(self, x, out=None):
        return torch.from_numpy(self.orc.posteriors(_np(x)))

    def loglik(self, x, prior, out=None):
        res = torch.from_numpy(self.orc.loglik(_np(x), _np(prior)))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def loglik_raw_rows(self, raw, utt_offsets, cmvn, feat_dim, context, prior, row_begin, rows, out):
        x = self._spliced(raw, utt_offsets, cmvn, feat_dim, context)[row_begin:row_begin + rows]
        out.copy_(torch.from_numpy(self.orc.loglik(x, _np(prior))))
        return out

    def halve_lr(self):
        self.calls.append("halve_lr")
        self.orc.halve_learning_rate()

    def set_active_layers(self, n):
        if n < 1 or n > self.num_layers:
            raise L.TfkError(L.TFK_EINVAL, "active layers %d" % n)
        self.orc.active = int(n)

    def set_dropout_seed(self, seed):
        self.drop_seed = int(seed)

    def kernel_launches(self):
        return 0

    def close(self):
        pass


class HostStager(object):
    """tfkaldi_b200.neuralNetworks.trainer._Stager without pinned memory, streams or a device"""

    def __init__(self, device, input_dim, capacity):
        self.capacity = capacity

    def stage(self, mats, targets):
        if isinstance(mats, (list, tuple)):
            x, y = np.concatenate(mats, axis=0), np.concatenate(targets).astype(np.int32)
            if x.shape[0] != y.shape[0]:
                raise ValueError("inputs hold %d frames but targets %d" % (x.shape[0], y.shape[0]))
        else:
            x, y = np.asarray(mats, np.float32), np.asarray(targets).astype(np.int32)
        if x.shape[0] > self.capacity:
            raise ValueError("micro-batch of %d frames exceeds the staging capacity %d" % (x.shape[0], self.capacity))
        return 0, torch.from_numpy(np.ascontiguousarray(x, np.float32)), torch.from_numpy(y)

    stage_pinned = stage

    def release(self, k):
        pass


class HostLane(object):
    """tfkaldi_b200.neuralNetworks.decoder._Lane without a device: plain host tensors, synchronous copies"""

    def __init__(self, device, tile, cols, slots):
        self.dev_out = [torch.empty((tile, cols), dtype=torch.float32) for _ in range(2)]
        self.pinned = [torch.empty((tile, cols), dtype=torch.float32) for _ in range(slots)]

    def before_compute(self, k):
        pass

    def to_host(self, k, p, n):
        self.pinned[p][:n].copy_(self.dev_out[k][:n])
        return lambda: None

    def upload(self, array, dtype):
        return torch.from_numpy(np.ascontiguousarray(array, dtype=dtype))
