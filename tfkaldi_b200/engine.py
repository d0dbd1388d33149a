"""Thin Python wrapper over the C-ABI handle.  PyTorch is used for device memory, streams and pinned
host buffers only; all arithmetic runs in libtfkaldi_b200.so (hand-written sm_100a CUDA)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class Engine:
    """One engine per GPU / process (rank).  Mirrors the `tf.Session` + graph pair of the reference
    (neuralNetworks/nnet.py:134, trainer.py:37-215, decoder.py:20-47)."""

    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, max_frames, *, nonlin="relu",
                 batch_norm=False, keep_prob=1.0, precision="bf16", device=None, seed=0, l2_norm=False):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise RuntimeError("tfkaldi_b200 needs a CUDA device (sm_100a); there is no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", int(device))
        cfg = L.TfkConfig()
        self.lib.tfk_default_config(C.byref(cfg))
        cfg.num_layers, cfg.input_dim, cfg.hidden_dim, cfg.output_dim = num_layers, input_dim, hidden_dim, output_dim
        cfg.max_frames = int(max_frames)
        codes = {"relu": L.TFK_NONLIN_RELU, "linear": L.TFK_NONLIN_LINEAR, "sigmoid": L.TFK_NONLIN_SIGMOID, "tanh": L.TFK_NONLIN_TANH}
        if nonlin not in codes:
            raise Exception("unkown nonlinearity")  # neuralNetworks/nnet.py:65 (sic)
        cfg.nonlin = codes[nonlin]
        cfg.batch_norm = 1 if batch_norm else 0
        cfg.keep_prob = float(keep_prob)
        cfg.precision = {"bf16": L.TFK_PREC_BF16, "bf16x3": L.TFK_PREC_BF16X3}[precision]
        cfg.device = self.device.index
        cfg.seed = int(seed)
        cfg.l2_norm = 1 if l2_norm else 0
        self.cfg = cfg
        self.precision = precision
        self.num_layers, self.input_dim, self.hidden_dim, self.output_dim = num_layers, input_dim, hidden_dim, output_dim
        self.max_frames = int(max_frames)
        self.batch_norm = bool(batch_norm)
        h = C.c_void_p()
        rc = self.lib.tfk_create(C.byref(cfg), C.byref(h))
        if rc != L.TFK_OK:
            raise L.TfkError(rc, self.lib.tfk_last_error(None).decode())
        self.h = h

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "h", None):
            self.lib.tfk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != L.TFK_OK:
            raise L.TfkError(rc, self.lib.tfk_last_error(self.h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def layer_shape(self, layer):
        k = self.input_dim if layer == 0 else self.hidden_dim
        n = self.hidden_dim if layer < self.num_layers else self.output_dim
        return k, n

    def _count(self, kind, layer):
        k, n = self.layer_shape(layer)
        return k * n if kind in (L.T_WEIGHTS, L.T_ADAM_M_W, L.T_ADAM_V_W, L.T_GRAD_W) else n

    # ------------------------------------------------------------------ tensors / scalars
    def set_tensor(self, kind, layer, array):
        a = np.ascontiguousarray(array, dtype=np.float32)
        if a.size != self._count(kind, layer):
            raise ValueError(f"tensor kind {kind} layer {layer}: expected {self._count(kind, layer)} elements, got {a.size}")
        self._check(self.lib.tfk_set_tensor(self.h, kind, layer, C.c_void_p(a.ctypes.data), a.size, self._stream()))

    def get_tensor(self, kind, layer):
        k, n = self.layer_shape(layer)
        shape = (k, n) if self._count(kind, layer) == k * n and kind in (L.T_WEIGHTS, L.T_ADAM_M_W, L.T_ADAM_V_W, L.T_GRAD_W) else (n,)
        out = np.empty(shape, dtype=np.float32)
        self._check(self.lib.tfk_get_tensor(self.h, kind, layer, C.c_void_p(out.ctypes.data), out.size, self._stream()))
        return out

    def set_scalar(self, kind, value):
        self._check(self.lib.tfk_set_scalar(self.h, kind, float(value)))

    def get_scalar(self, kind):
        v = C.c_double()
        self._check(self.lib.tfk_get_scalar(self.h, kind, C.byref(v), self._stream()))
        return v.value

    def load_params(self, params: dict):
        """params: {'W{l}', 'b{l}', 'beta{l}', 'moving_mean{l}', 'moving_var{l}'} numpy arrays."""
        names = {"W": L.T_WEIGHTS, "b": L.T_BIASES, "beta": L.T_BN_BETA, "moving_mean": L.T_BN_MOVING_MEAN,
                 "moving_var": L.T_BN_MOVING_VAR}
        for key, val in params.items():
            stem = key.rstrip("0123456789")
            self.set_tensor(names[stem], int(key[len(stem):]), val)

    def dump_params(self) -> dict:
        out = {}
        for l in range(self.num_layers + 1):
            out[f"W{l}"] = self.get_tensor(L.T_WEIGHTS, l)
            out[f"b{l}"] = self.get_tensor(L.T_BIASES, l)
            if self.batch_norm and l < self.num_layers:
                out[f"beta{l}"] = self.get_tensor(L.T_BN_BETA, l)
                out[f"moving_mean{l}"] = self.get_tensor(L.T_BN_MOVING_MEAN, l)
                out[f"moving_var{l}"] = self.get_tensor(L.T_BN_MOVING_VAR, l)
        return out

    # ------------------------------------------------------------------ device-side steps
    def _dev_f32(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return x.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()

    def _dev_i32(self, y):
        if isinstance(y, np.ndarray):
            y = torch.from_numpy(np.ascontiguousarray(y.astype(np.int32)))
        return y.to(self.device, dtype=torch.int32, non_blocking=True).contiguous()

    def accumulate(self, x, labels):
        x, labels = self._dev_f32(x), self._dev_i32(labels)
        self._check(self.lib.tfk_accumulate(self.h, _ptr(x), _ptr(labels), x.shape[0], self._stream()))

    def accumulate_raw(self, raw, utt_offsets, cmvn, labels, feat_dim, context):
        """device-side CMVN + splice feeder: raw [R, D] fp32, utt_offsets int32 [U+1], cmvn [U, 2, D] (mean, 1/std)"""
        raw, cmvn = self._dev_f32(raw), self._dev_f32(cmvn)
        utt_offsets, labels = self._dev_i32(utt_offsets), self._dev_i32(labels)
        self._check(self.lib.tfk_accumulate_raw(self.h, _ptr(raw), _ptr(utt_offsets), utt_offsets.shape[0] - 1, _ptr(cmvn),
                                                _ptr(labels), raw.shape[0], int(feat_dim), int(context), self._stream()))

    def loglik_raw(self, raw, utt_offsets, cmvn, feat_dim, context, prior, out=None):
        raw, cmvn, prior = self._dev_f32(raw), self._dev_f32(cmvn), self._dev_f32(prior)
        utt_offsets = self._dev_i32(utt_offsets)
        if out is None:
            out = torch.empty((raw.shape[0], self.output_dim), dtype=torch.float32, device=self.device)
        self._check(self.lib.tfk_forward_loglik_raw(self.h, _ptr(raw), _ptr(utt_offsets), utt_offsets.shape[0] - 1, _ptr(cmvn),
                                                    raw.shape[0], int(feat_dim), int(context), _ptr(prior), _ptr(out), self._stream()))
        return out

    def loglik_raw_rows(self, raw, utt_offsets, cmvn, feat_dim, context, prior, row_begin, rows, out):
        """rows [row_begin, row_begin + rows) of loglik_raw; raw / utt_offsets / cmvn / prior / out are DEVICE tensors"""
        self._check(self.lib.tfk_forward_loglik_raw_rows(self.h, _ptr(raw), _ptr(utt_offsets), utt_offsets.shape[0] - 1, _ptr(cmvn),
                                                         int(row_begin), int(rows), int(feat_dim), int(context), _ptr(prior), _ptr(out),
                                                         self._stream()))
        return out

    def apply(self, lr, want_loss=True):
        if want_loss:
            out = C.c_float()
            self._check(self.lib.tfk_apply(self.h, float(lr), C.byref(out), self._stream()))
            return out.value
        self._check(self.lib.tfk_apply(self.h, float(lr), None, self._stream()))
        return None

    def train_step(self, x, labels, lr, want_loss=True):
        """accumulate + apply for one micro-batch with the Adam pass overlapped (tfk_train_step)"""
        x, labels = self._dev_f32(x), self._dev_i32(labels)
        out = C.c_float()
        self._check(self.lib.tfk_train_step(self.h, _ptr(x), _ptr(labels), x.shape[0], float(lr),
                                            C.byref(out) if want_loss else None, self._stream()))
        return out.value if want_loss else None

    def train_step_raw(self, raw, utt_offsets, cmvn, labels, feat_dim, context, lr, want_loss=True):
        """train_step fed by the device-side CMVN + splice feeder (tfk_train_step_raw); arguments as accumulate_raw"""
        raw, cmvn = self._dev_f32(raw), self._dev_f32(cmvn)
        utt_offsets, labels = self._dev_i32(utt_offsets), self._dev_i32(labels)
        out = C.c_float()
        self._check(self.lib.tfk_train_step_raw(self.h, _ptr(raw), _ptr(utt_offsets), utt_offsets.shape[0] - 1, _ptr(cmvn),
                                                _ptr(labels), raw.shape[0], int(feat_dim), int(context), float(lr),
                                                C.byref(out) if want_loss else None, self._stream()))
        return out.value if want_loss else None

    def last_loss(self):
        """the mean loss of the last step run with want_loss=False (blocks until that step has finished)"""
        out = C.c_float()
        self._check(self.lib.tfk_last_loss(self.h, C.byref(out), self._stream()))
        return out.value

    def eval_accumulate(self, x, labels):
        x, labels = self._dev_f32(x), self._dev_i32(labels)
        self._check(self.lib.tfk_eval_accumulate(self.h, _ptr(x), _ptr(labels), x.shape[0], self._stream()))

    def eval_finish(self):
        out = C.c_float()
        self._check(self.lib.tfk_eval_finish(self.h, C.byref(out), self._stream()))
        return out.value

    def posteriors(self, x, out=None):
        x = self._dev_f32(x)
        if out is None:
            out = torch.empty((x.shape[0], self.output_dim), dtype=torch.float32, device=self.device)
        self._check(self.lib.tfk_forward_posteriors(self.h, _ptr(x), x.shape[0], _ptr(out), self._stream()))
        return out

    def loglik(self, x, prior, out=None):
        x, prior = self._dev_f32(x), self._dev_f32(prior)
        if out is None:
            out = torch.empty((x.shape[0], self.output_dim), dtype=torch.float32, device=self.device)
        self._check(self.lib.tfk_forward_loglik(self.h, _ptr(x), x.shape[0], _ptr(prior), _ptr(out), self._stream()))
        return out

    def fflayer_fwd(self, layer, x, training=False):
        x = self._dev_f32(x)
        _, n = self.layer_shape(layer)
        y = torch.empty((x.shape[0], n), dtype=torch.float32, device=self.device)
        self._check(self.lib.tfk_fflayer_fwd(self.h, layer, _ptr(x), _ptr(y), x.shape[0], 1 if training else 0, self._stream()))
        return y

    def fflayer_bwd(self, layer, dy, want_dx=True):
        dy = self._dev_f32(dy)
        k, _ = self.layer_shape(layer)
        dx = torch.empty((dy.shape[0], k), dtype=torch.float32, device=self.device) if want_dx else None
        self._check(self.lib.tfk_fflayer_bwd(self.h, layer, _ptr(dy), _ptr(dx), dy.shape[0], self._stream()))
        return dx

    def softmax_ce(self, logits, labels, want_grad=True):
        logits, labels = self._dev_f32(logits), self._dev_i32(labels)
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        d = torch.empty_like(logits) if want_grad else None
        self._check(self.lib.tfk_softmax_ce(self.h, _ptr(logits), _ptr(labels), logits.shape[0], _ptr(loss), _ptr(d), self._stream()))
        return loss, d

    def halve_lr(self):
        self._check(self.lib.tfk_halve_lr(self.h))

    def set_active_layers(self, n):
        self._check(self.lib.tfk_set_active_layers(self.h, int(n)))

    def set_dropout_seed(self, seed):
        self._check(self.lib.tfk_set_dropout_seed(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def activation(self, layer, frames):
        """stored output of hidden layer `layer` from the last forward pass, fp32 [frames, hidden_dim] (diagnostic)"""
        out = torch.empty((int(frames), self.hidden_dim), dtype=torch.float32, device=self.device)
        self._check(self.lib.tfk_get_activation(self.h, int(layer), _ptr(out), int(frames), self._stream()))
        return out

    # ------------------------------------------------------------------ data parallel
    def init_comm_from_torch(self):
        """Create the engine's own NCCL communicator, exchanging the unique id over torch.distributed."""
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            rc = self.lib.tfk_comm_unique_id(ident)
            if rc != L.TFK_OK:
                raise L.TfkError(rc, self.lib.tfk_last_error(None).decode())
        payload = [bytes(ident)]
        dist.broadcast_object_list(payload, src=0)
        ident = (C.c_uint8 * 128).from_buffer_copy(payload[0])
        self._check(self.lib.tfk_comm_init(self.h, ident, rank, world))
        mode = os.environ.get("TFK_DP_MODE", "")
        if mode != "allreduce" and world & (world - 1) == 0:
            # peer memory needs every rank on ONE host with P2P access between all device pairs; otherwise the
            # NCCL transports (reduce-scatter / all-gather inside tfk_apply) stay in use
            import socket

            where = [None] * world
            dist.all_gather_object(where, (socket.gethostname(), self.device.index))
            same_host = len({h for h, _ in where}) == 1
            p2p = same_host and all(d == self.device.index or torch.cuda.can_device_access_peer(self.device.index, d) for _, d in where)
            flags = [None] * world
            dist.all_gather_object(flags, bool(p2p))
            if not all(flags):
                return
            # single-node peer memory: let the wgrad epilogues reduce-add into the owners' accumulators
            mine = (C.c_uint8 * 256)()
            self._check(self.lib.tfk_ipc_export(self.h, mine))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(mine))
            blob = (C.c_uint8 * (256 * world)).from_buffer_copy(b"".join(handles))
            self._check(self.lib.tfk_ipc_import(self.h, blob, world))
            dist.barrier()

    # ------------------------------------------------------------------ measurement
    def enable_timers(self, on=True):
        self._check(self.lib.tfk_enable_timers(self.h, 1 if on else 0))

    def timers(self):
        ms = (C.c_double * L.NUM_TIMERS)()
        n = (C.c_int64 * L.NUM_TIMERS)()
        self._check(self.lib.tfk_get_timers(self.h, ms, n))
        return {name: (ms[i], n[i]) for i, name in enumerate(L.TIMER_NAMES)}

    def kernel_launches(self):
        return int(self.lib.tfk_kernel_launches(self.h))
