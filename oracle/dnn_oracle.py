"""CPU oracle for the DNN / CrossEnthropyTrainer / Decoder hot path of vrenkens/tfkaldi.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product (tfkaldi_b200/) never does.

PARITY STATUS: *unpinned by the reference*.  The reference holds no tests, fixtures or golden
vectors, is Python-2 / TensorFlow-0.1x only and cannot run in this image.  The
arithmetic below is an explicit fp32 numpy restatement (forward, backward and optimizer written out,
no autograd) of the reference call sites cited at each function, with TensorFlow r0.11/r0.12 default
semantics restated from SURVEY.md Appendix A.  What pins it instead (tests/test_oracle_*.py):
hand-derived known-answer tests, the structural facts the reference code implies (zero-initialised
output layer => first loss == ln O; first step moves only the output layer), and an independent
float64 torch-autograd cross-check of every gradient.

All reference paths are relative to /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .philox import dropout_keep_mask

F32 = np.float32

_MATMUL_BACKEND = "numpy"


def set_matmul_backend(name: str) -> None:
    """'numpy' (OpenBLAS) or 'torch' (MKL/oneDNN, all host threads) for the fp32 GEMMs."""
    global _MATMUL_BACKEND
    if name not in ("numpy", "torch"):
        raise ValueError(name)
    _MATMUL_BACKEND = name


def _mm(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    if _MATMUL_BACKEND == "torch":
        import torch

        # from_numpy keeps transposed views as strides: MKL takes them as transposition flags, no copy
        return torch.mm(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    return np.matmul(a, b)


@dataclass
class OracleConfig:
    num_layers: int  # hidden layers               classifiers/dnn.py:13
    input_dim: int  # spliced feature dimension   nnet.py:40
    hidden_dim: int
    output_dim: int
    nonlin: str = "relu"  # 'relu' | 'sigmoid' | 'tanh' | 'linear'   nnet.py:47-62
    batch_norm: bool = False  # nnet.py:42
    l2_norm: bool = False  # nnet.py:67-68
    keep_prob: float = 1.0  # conf['dropout'] = KEEP prob nnet.py:70-72, activation.py:141
    bn_eps: float = 1e-3  # tf.contrib.layers.batch_norm defaults (App. A.3)
    bn_decay: float = 0.999
    adam_beta1: float = 0.9  # tf.train.AdamOptimizer defaults, trainer.py:115
    adam_beta2: float = 0.999
    adam_eps: float = 1e-8


def reference_init(cfg: OracleConfig, rng: np.random.Generator) -> dict:
    """Initial parameters as the reference draws them.

    classifiers/layer.py:39-48: hidden W ~ N(0, (1/sqrt(fan_in))^2), b = 0;
    classifiers/dnn.py:67-68: the output FFLayer gets weights_std=0 -> W == 0 exactly.
    BN: beta = 0, moving_mean = 0, moving_variance = 1 (tf.contrib.layers.batch_norm).
    """
    p = {}
    for l in range(cfg.num_layers + 1):
        k = cfg.input_dim if l == 0 else cfg.hidden_dim
        n = cfg.hidden_dim if l < cfg.num_layers else cfg.output_dim
        if l < cfg.num_layers:
            p[f"W{l}"] = (rng.standard_normal((k, n)) / math.sqrt(k)).astype(F32, copy=False)
        else:
            p[f"W{l}"] = np.zeros((k, n), F32)
        p[f"b{l}"] = np.zeros(n, F32)
        if cfg.batch_norm and l < cfg.num_layers:
            p[f"beta{l}"] = np.zeros(n, F32)
            p[f"moving_mean{l}"] = np.zeros(n, F32)
            p[f"moving_var{l}"] = np.ones(n, F32)
    return p


@dataclass
class _LayerCache:
    x: np.ndarray  # layer input
    z: np.ndarray | None = None  # linear output (kept for BN backward)
    xhat: np.ndarray | None = None
    rstd: np.ndarray | None = None
    y: np.ndarray | None = None  # output after the whole activation chain
    a: np.ndarray | None = None  # output of the nonlinearity (before dropout)
    keep: np.ndarray | None = None  # dropout keep mask (training with keep_prob < 1)
    sig: np.ndarray | None = None  # L2Norm: per-frame mean square of a


class OracleDNN:
    """Explicit fp32 restatement of DNN + Trainer + Decoder state and steps."""

    def __init__(self, cfg: OracleConfig, params: dict):
        self.cfg = cfg
        self.L = cfg.num_layers
        self.active = cfg.num_layers  # classifiers/dnn.py:85-104 layer-wise growth ("initialisedlayers"+1)
        self.p = {k: np.array(v, dtype=F32, copy=True) for k, v in params.items()}
        # trainable variables: weights, biases and (batch_norm) beta of every layer   trainer.py:82
        self.trainable = [k for k in self.p if not k.startswith("moving")]
        self.m = {k: np.zeros_like(self.p[k]) for k in self.trainable}  # Adam slots
        self.v = {k: np.zeros_like(self.p[k]) for k in self.trainable}
        self.grads = {k: np.zeros_like(self.p[k]) for k in self.trainable}  # trainer.py:118-122
        self.loss_sum = 0.0  # batch_loss        trainer.py:91-93
        self.num_frames = 0  # train/num_frames  trainer.py:126-128
        self.global_step = 0  # trainer.py:98-100
        # Adam's own step count: tf.train.AdamOptimizer keeps beta1_power / beta2_power as optimizer variables
        # (trainer.py:115); they are initialised by init_op and are NOT in the `train_variables` saver
        # (trainer.py:204-205), so restore_trainer rewinds global_step but never the beta powers
        self.adam_step = 0
        self.lr_fact = 1.0  # trainer.py:104-106

    # ------------------------------------------------------------------ forward
    def _nonlin(self, a):
        if self.cfg.nonlin == "relu":
            return np.maximum(a, F32(0))
        if self.cfg.nonlin == "linear":
            return a
        if self.cfg.nonlin == "sigmoid":
            return (F32(1) / (F32(1) + np.exp(-a))).astype(F32, copy=False)
        if self.cfg.nonlin == "tanh":
            return np.tanh(a).astype(F32, copy=False)
        raise Exception("unkown nonlinearity")  # nnet.py:65 (sic)

    def forward(self, x: np.ndarray, training: bool, dropout_seed: int = 0):
        """DNN.__call__ on packed frames (classifiers/dnn.py:73-109): returns (logits, caches).

        Per hidden layer (classifiers/layer.py:52-56 + activation chain order nnet.py:42-72):
            z = x W + b -> [batch_norm] -> nonlin -> [dropout (training only)]
        Output layer: identity activation only (dnn.py:67-68).
        """
        cfg = self.cfg
        a = np.ascontiguousarray(x, dtype=F32)
        caches = []
        for l in range(self.active):
            c = _LayerCache(x=a)
            z = _mm(a, self.p[f"W{l}"]) + self.p[f"b{l}"]
            if cfg.batch_norm:
                c.z = z
                if training:
                    # tf.nn.moments over the micro-batch: biased variance (App. A.3)
                    mu = z.mean(axis=0, dtype=np.float64).astype(F32, copy=False)
                    var = np.mean(np.square(z - mu, dtype=np.float64), axis=0).astype(F32, copy=False)
                    # assign_moving_average, run through UPDATE_OPS once per micro-batch (trainer.py:164-168)
                    d = F32(1.0 - cfg.bn_decay)
                    self.p[f"moving_mean{l}"] -= d * (self.p[f"moving_mean{l}"] - mu)
                    self.p[f"moving_var{l}"] -= d * (self.p[f"moving_var{l}"] - var)
                else:
                    mu, var = self.p[f"moving_mean{l}"], self.p[f"moving_var{l}"]
                c.rstd = (F32(1.0) / np.sqrt(var + F32(cfg.bn_eps))).astype(F32, copy=False)
                c.xhat = ((z - mu) * c.rstd).astype(F32, copy=False)
                h = c.xhat + self.p[f"beta{l}"]  # center=True, scale=False
            else:
                h = z
            h = self._nonlin(h)
            c.a = h
            if cfg.l2_norm:
                # L2Norm._apply_func (classifiers/activation.py:103-111): sig = mean over columns of a^2;
                # a / sig where sig > 1, else a  (the mean SQUARE, not its root - reference quirk kept)
                c.sig = np.mean(np.square(h, dtype=np.float64), axis=1, keepdims=True).astype(F32)
                h = np.where(c.sig > 1, h / c.sig, h).astype(F32, copy=False)
            if training and cfg.keep_prob < 1.0:
                # tf.nn.dropout: x / keep * floor(keep + u)   classifiers/activation.py:140-141
                c.keep = dropout_keep_mask(dropout_seed + l, h.shape[0], h.shape[1], cfg.keep_prob)
                h = np.where(c.keep, h * F32(1.0 / cfg.keep_prob), F32(0)).astype(F32, copy=False)
            c.y = h.astype(F32, copy=False)
            caches.append(c)
            a = c.y
        cl = _LayerCache(x=a)
        logits = (_mm(a, self.p[f"W{self.L}"]) + self.p[f"b{self.L}"]).astype(F32, copy=False)
        caches.append(cl)
        return logits, caches

    # ------------------------------------------------------------------ loss
    @staticmethod
    def softmax_ce(logits: np.ndarray, labels: np.ndarray):
        """CrossEnthropyTrainer.compute_loss (trainer.py:514-531): SUM over frames of
        softmax_cross_entropy_with_logits(logits, one_hot(labels)); returns (loss_sum, dlogits)
        with dlogits = softmax - onehot (not divided by the frame count).
        Labels outside [0,O) give an all-zero one-hot row (tf.one_hot) => loss 0; the gradient of that row is still
        softmax - onehot = softmax (the TF op computes backprop = prob - labels whatever the labels sum to).  Cannot
        occur through the reference's data path (alignments are < num_labels), kept for literal fidelity."""
        z = logits.astype(F32, copy=False)
        mx = z.max(axis=1, keepdims=True)
        e = np.exp(z - mx)
        s = e.sum(axis=1, keepdims=True, dtype=F32)
        labels = np.asarray(labels).astype(np.int64)
        ok = (labels >= 0) & (labels < z.shape[1])
        idx = np.where(ok, labels, 0)
        rows = np.arange(z.shape[0])
        row_loss = (np.log(s[:, 0]) + mx[:, 0] - z[rows, idx]).astype(F32, copy=False)
        row_loss = np.where(ok, row_loss, F32(0))
        d = (e / s).astype(F32, copy=False)
        d[rows[ok], labels[ok]] -= F32(1)
        return float(row_loss.sum(dtype=np.float64)), d.astype(F32, copy=False)

    # ------------------------------------------------------------------ backward
    def backward(self, caches, dlogits: np.ndarray, relu_pass=None) -> dict:
        """tf.gradients(loss, params) (trainer.py:155), written out.  Returns {name: grad}.

        relu_pass (test aid): optional list of boolean [B, N] arrays, one per hidden layer, used INSTEAD of this
        implementation's own `a > 0` as the ReLU gradient gate (activation.py:84).  A pre-activation within rounding
        distance of zero lands on either side in any two fp32 implementations; replaying the backward pass with the
        other implementation's activation pattern takes that discontinuity out of a gradient comparison."""
        cfg = self.cfg
        g = {}
        L = self.L
        cl = caches[-1]
        g[f"W{L}"] = _mm(cl.x.T, dlogits).astype(F32, copy=False)
        g[f"b{L}"] = dlogits.sum(axis=0, dtype=F32)
        da = _mm(dlogits, self.p[f"W{L}"].T).astype(F32, copy=False)
        for l in range(self.active - 1, -1, -1):
            c = caches[l]
            # dropout backward, then the nonlinearity's slope at its own output a = f(h)
            dh = da
            if c.keep is not None:
                dh = np.multiply(dh * F32(1.0 / cfg.keep_prob), c.keep, dtype=F32)
            if cfg.l2_norm:
                # y = a / sig(a): dy/da = I/sig - a (2a/N)^T / sig^2 on the normalised rows, identity elsewhere
                n = F32(c.a.shape[1])
                dot = np.sum(dh * c.a, axis=1, keepdims=True, dtype=np.float64).astype(F32)
                dn = dh / c.sig - c.a * (F32(2) * dot / (n * c.sig * c.sig))
                dh = np.where(c.sig > 1, dn, dh).astype(F32, copy=False)
            if cfg.nonlin == "relu":
                dh = np.multiply(dh, (c.a > 0) if relu_pass is None else relu_pass[l], dtype=F32)
            elif cfg.nonlin == "sigmoid":
                dh = (dh * (c.a * (F32(1) - c.a))).astype(F32, copy=False)
            elif cfg.nonlin == "tanh":
                dh = (dh * (F32(1) - c.a * c.a)).astype(F32, copy=False)
            if cfg.batch_norm:
                # y = xhat + beta: dbeta = sum dy; dz = r * (dy - mean(dy) - xhat * mean(dy*xhat))  (App. A.3)
                g[f"beta{l}"] = dh.sum(axis=0, dtype=F32)
                m1 = dh.mean(axis=0, dtype=np.float64).astype(F32, copy=False)
                m2 = (dh * c.xhat).mean(axis=0, dtype=np.float64).astype(F32, copy=False)
                dz = (c.rstd * (dh - m1 - c.xhat * m2)).astype(F32, copy=False)
            else:
                dz = dh
            g[f"W{l}"] = _mm(c.x.T, dz).astype(F32, copy=False)
            if cfg.batch_norm:
                # z = xW + b enters a batch-normalised layer only through z - mean_B(z): sum_B dz == 0 identically
                # (sum xhat = 0).  TensorFlow's fp32 graph evaluates this zero as round-off noise, which Adam then
                # normalises into a random walk of a parameter the training output does not depend on; that noise is
                # not reproducible by any second implementation, so the oracle (and the engine) use the exact value.
                g[f"b{l}"] = np.zeros(dz.shape[1], F32)
            else:
                g[f"b{l}"] = dz.sum(axis=0, dtype=F32)
            if l > 0:
                da = _mm(dz, self.p[f"W{l}"].T).astype(F32, copy=False)
        return g

    # ------------------------------------------------------------------ trainer steps
    def accumulate(self, x, labels, dropout_seed: int = 0, relu_pass=None) -> float:
        """== update_gradients_op.run(feed) (trainer.py:165-169, 328): one micro-batch."""
        logits, caches = self.forward(x, training=True, dropout_seed=dropout_seed)
        loss, d = self.softmax_ce(logits, labels)
        g = self.backward(caches, d, relu_pass=relu_pass)
        for k, v in g.items():
            self.grads[k] += v  # grads[p].assign_add(batchgrads[p])
        self.loss_sum += loss  # batch_loss.assign_add(loss)
        self.num_frames += int(x.shape[0])  # num_frames.assign_add(sum(target_seq_length))
        return loss

    def apply(self, lr: float) -> float:
        """== run([average_loss, apply_gradients_op]) + re-initialisers (trainer.py:174-184, 337-352).

        meangrad = clip(grad / num_frames, -1, 1); TF ApplyAdam:
            t += 1; lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
            var -= lr_t*m/(sqrt(v)+eps)           (eps OUTSIDE the bias correction, App. A.8)
        `lr` is the decayed rate lr0*decay^(global_step/num_steps) (trainer.py:110-112); multiplied here
        by learning_rate_fact.  Returns batch_loss/num_frames evaluated with the pre-update weights."""
        cfg = self.cfg
        mean_loss = self.loss_sum / float(self.num_frames)
        self.global_step += 1
        self.adam_step += 1
        t = float(self.adam_step)
        b1, b2 = cfg.adam_beta1, cfg.adam_beta2
        lr_t = F32(lr * self.lr_fact * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
        nf = F32(self.num_frames)
        for k in self.trainable:
            if _MATMUL_BACKEND == "torch":
                # same fp32 formulas through torch's multi-threaded element-wise kernels (in place on
                # the numpy buffers) so the timed CPU baseline uses every host core, as TF-CPU's Eigen would
                import torch

                g, m, v, w = (torch.from_numpy(a) for a in (self.grads[k], self.m[k], self.v[k], self.p[k]))
                ghat = torch.clamp(g / float(nf), -1.0, 1.0)
                m.add_((ghat - m) * float(F32(1.0 - b1)))
                v.add_((ghat * ghat - v) * float(F32(1.0 - b2)))
                w.sub_((m * float(lr_t)) / (torch.sqrt(v) + float(F32(cfg.adam_eps))))
                g.zero_()
                continue
            ghat = np.clip(self.grads[k] / nf, F32(-1), F32(1)).astype(F32, copy=False)
            self.m[k] += (ghat - self.m[k]) * F32(1.0 - b1)
            self.v[k] += (ghat * ghat - self.v[k]) * F32(1.0 - b2)
            self.p[k] -= (self.m[k] * lr_t) / (np.sqrt(self.v[k]) + F32(cfg.adam_eps))
            self.grads[k][...] = 0  # init_grads
        self.loss_sum, self.num_frames = 0.0, 0  # init_loss, init_num_frames
        return mean_loss

    def eval_accumulate(self, x, labels) -> float:
        """== update_valid_loss.run(feed) (trainer.py:186-195, 428): eval tower + CE."""
        logits, _ = self.forward(x, training=False)
        loss, _ = self.softmax_ce(logits, labels)
        self.loss_sum += loss
        self.num_frames += int(x.shape[0])
        return loss

    def eval_finish(self) -> float:
        """== average_loss.eval() + init_loss/init_num_frames (trainer.py:435-439)."""
        out = self.loss_sum / float(self.num_frames)
        self.loss_sum, self.num_frames = 0.0, 0
        return out

    def halve_learning_rate(self):  # trainer.py:141-142
        self.lr_fact /= 2.0

    # ------------------------------------------------------------------ decoder
    def posteriors(self, x) -> np.ndarray:
        """Decoder.__call__ (decoder.py:44, 49-71): eval-mode forward + softmax."""
        logits, _ = self.forward(x, training=False)
        mx = logits.max(axis=1, keepdims=True)
        e = np.exp(logits - mx)
        return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32, copy=False)

    def loglik(self, x, prior) -> np.ndarray:
        """Nnet.decode post-processing (nnet.py:280-286): log(posterior / prior), float32, NO flooring
        (the reference's np.where result is discarded)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.log(self.posteriors(x) / np.asarray(prior, dtype=F32)).astype(F32, copy=False)


def learning_rate(lr0: float, decay: float, global_step: int, num_steps: int) -> float:
    """tf.train.exponential_decay(lr0, global_step, num_steps, decay), non-staircase (trainer.py:110-112)."""
    return lr0 * decay ** (float(global_step) / float(num_steps))


def compute_prior(target_arrays, num_labels: int) -> np.ndarray:
    """nnet.py:241-244 + batchdispenser.py:128-145: bincount over ALL targets, float32, normalised."""
    count = np.bincount(np.concatenate([np.asarray(t) for t in target_arrays]), minlength=num_labels)
    prior = count.astype(F32, copy=False)
    return prior / prior.sum()
