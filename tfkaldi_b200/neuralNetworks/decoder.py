"""Decoding environment with the reference's Decoder API (reference: neuralNetworks/decoder.py).

`Decoder(classifier, input_dim, max_length)(inputs)` returned softmax posteriors from
`outputs.eval(feed)` on the utterance padded to max_length (decoder.py:49-71); here it is one
tfk_forward_posteriors call on the unpadded frames.  `loglik` additionally fuses Nnet.decode's
host-side `np.log(output / prior)` (nnet.py:280-286) into the output kernel."""
import collections
import queue
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from ..engine import Engine
from ..processing.feeder import cmvn_coefficients


class Decoder(object):
    def __init__(self, classifier, input_dim, max_length, *, precision="bf16x3", device=None, max_frames=16384):
        """precision defaults to the fp32-equivalent mode: posteriors / log-likelihoods are what the 1e-3 contract
        against the reference's fp32 TensorFlow arithmetic is stated on (decoder.py:26-27, 44); "bf16" is 3x faster
        and measured at full size in tests/test_gpu_zfullsize.py (log-likelihood error ~1e-2 absolute)."""
        spec = classifier.engine_spec(input_dim)
        self.max_length = max_length
        self.input_dim = input_dim
        self.engine = Engine(spec["num_layers"], spec["input_dim"], spec["hidden_dim"], spec["output_dim"],
                             min(int(max_frames), max(int(max_length), 1)), nonlin=spec["nonlin"],
                             batch_norm=spec["batch_norm"], keep_prob=spec["keep_prob"], l2_norm=spec["l2_norm"], precision=precision, device=device)
        self._trainer_like = None

    def __call__(self, inputs):
        """[N, F] numpy -> [N, O] numpy posteriors (eval mode: moving-stat batch norm, no dropout)"""
        return self.engine.posteriors(np.ascontiguousarray(inputs, dtype=np.float32)).cpu().numpy()

    def loglik(self, inputs, prior, out=None):
        """device tensor [N, O] = log(softmax / prior), no flooring (nnet.py:280-286)"""
        return self.engine.loglik(inputs, prior, out=out)

    def restore(self, filename):
        """load the model written by Trainer.save_model — or by the reference's own saver (decoder.py:73-81)"""
        from .trainer import MODEL_NAMES, read_model_file

        params = {}
        for key, val in read_model_file(filename).items():
            if key == "Classifier/initialisedlayers":
                self.engine.set_active_layers(int(val) + 1)
                continue
            if not key.startswith("Classifier/"):
                continue
            _, layer, rest = key.split("/", 2)
            params[MODEL_NAMES[rest] + layer[len("layer"):]] = val
        self.engine.load_params(params)


class _Lane(object):
    """the device-side plumbing of LoglikStreamer: output tiles on the device, a copy stream, pinned host slots"""

    def __init__(self, device, tile, cols, slots):
        self.device = device
        self.dev_out = [torch.empty((tile, cols), dtype=torch.float32, device=device) for _ in range(2)]
        self.pinned = [torch.empty((tile, cols), dtype=torch.float32, pin_memory=True) for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.dev_free = [None, None]  # event: the copy out of dev_out[k] has finished

    def before_compute(self, k):
        if self.dev_free[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.dev_free[k])

    def to_host(self, k, p, n):
        """queue the copy of dev_out[k][:n] into pinned slot p behind the compute stream; returns a wait() callable"""
        done = torch.cuda.Event()
        computed = torch.cuda.Event()
        computed.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(computed)
            self.pinned[p][:n].copy_(self.dev_out[k][:n], non_blocking=True)
            done.record(self.copy_stream)
        self.dev_free[k] = done
        return done.synchronize

    def upload(self, array, dtype):
        a = np.ascontiguousarray(array, dtype=dtype)
        if not a.flags.writeable:  # views into a memory-mapped archive are read-only
            a = a.copy()
        return torch.from_numpy(a).to(self.device, non_blocking=True)


class LoglikStreamer(object):
    """Decoder -> ArkWriter pipeline for Nnet.decode (reference: neuralNetworks/nnet.py:267-289, which evaluates one
    utterance, divides by the prior and takes the log on the host, and appends the matrix to the archive, one after the
    other).  Here an utterance is cut into tiles of `tile` frames; while tile i+1 runs through the network, tile i
    travels device -> pinned host memory on a copy stream and tile i-1 is written into the archive by a pool of writer
    threads through a shared mapping of the archive (ArkWriter.begin_utt / write_rows / finish_utt: the same bytes write_next_utt
    produces).  The next utterance starts while the last tiles of the previous one are still being written."""

    def __init__(self, decoder, writer, prior, tile=None, slots=4, io_threads=8, lane=None):
        self.decoder, self.writer = decoder, writer
        eng = decoder.engine
        self.tile = int(tile or eng.max_frames)
        self.cols = eng.output_dim
        self.lane = lane if lane is not None else _Lane(eng.device, self.tile, self.cols, slots)
        self.prior = self.lane.upload(prior, np.float32)
        self.free = queue.Queue()
        for p in range(len(self.lane.pinned)):
            self.free.put(p)
        self.io_threads = io_threads
        self.pool = ThreadPoolExecutor(max_workers=io_threads)
        self.pending = collections.deque()  # (entry, [futures]) in utterance order
        self.turn = 0

    # -- writer side
    def _write_part(self, wait, entry, first_row, p, lo, hi, remaining):
        wait()
        block = self.lane.pinned[p][lo:hi].numpy()
        self.writer.write_rows(entry, first_row + lo, block)
        with remaining[1]:
            remaining[0] -= 1
            last = remaining[0] == 0
        if last:
            self.free.put(p)  # the slot may be overwritten by the next copy

    def _emit(self, entry, first_row, k, n, futures):
        import threading

        p = self.free.get()  # back-pressure: blocks while every pinned slot is still being written
        wait = self.lane.to_host(k, p, n)
        parts = max(1, min(self.io_threads, n // 512))
        remaining = [parts, threading.Lock()]
        step = -(-n // parts)
        for i in range(parts):
            lo, hi = i * step, min(n, (i + 1) * step)
            futures.append(self.pool.submit(self._write_part, wait, entry, first_row, p, lo, hi, remaining))

    def _retire(self, block):
        while self.pending and (block or all(f.done() for f in self.pending[0][1])):
            entry, futures = self.pending.popleft()
            for f in futures:
                f.result()  # re-raises a writer thread's exception
            self.writer.finish_utt(entry)

    # -- producer side
    def _run(self, utt_id, frames, compute):
        self._retire(block=False)
        entry = self.writer.begin_utt(utt_id, frames, self.cols)
        futures = []
        for t0 in range(0, frames, self.tile):
            n = min(self.tile, frames - t0)
            k = self.turn
            self.turn ^= 1
            self.lane.before_compute(k)
            compute(t0, n, self.lane.dev_out[k][:n])
            self._emit(entry, t0, k, n, futures)
        self.pending.append((entry, futures))

    def decode_spliced(self, utt_id, utt_mat):
        """utt_mat: [T, input_dim] normalised + spliced features (FeatureReader.get_utt)"""
        eng = self.decoder.engine
        self._run(utt_id, utt_mat.shape[0], lambda t0, n, out: eng.loglik(utt_mat[t0:t0 + n], self.prior, out=out))

    def decode_raw(self, utt_id, raw, cmvn_stats, context_width):
        """raw: [T, D] un-normalised features, cmvn_stats: the speaker's accumulated statistics [2, D+1]; CMVN and the
        splice (feature_reader.py:42-60) run on the device: 2k+1 times fewer bytes cross PCIe on the way in"""
        eng = self.decoder.engine
        frames, dim = raw.shape
        d_raw = self.lane.upload(raw, np.float32)
        d_off = self.lane.upload(np.array([0, frames]), np.int32)
        d_cmvn = self.lane.upload(cmvn_coefficients(cmvn_stats)[None], np.float32)
        self._run(utt_id, frames, lambda t0, n, out: eng.loglik_raw_rows(d_raw, d_off, d_cmvn, dim, context_width, self.prior, t0, n, out))

    def close(self):
        self._retire(block=True)
        self.pool.shutdown(wait=True)
