"""Prefetching feeder: a background thread turns the dispenser's utterances into packed RAW batches in pinned
host memory; the training thread copies a ready batch to the device on a side stream while the previous step is
still computing (BASELINE.json north_star: "processing.batchdispenser/ArkReader feeds pinned host buffers copied on
a side stream").

What it replaces in the reference: `dispenser.get_batch()` called synchronously between two `trainer.update()`s
(neuralNetworks/nnet.py:157-160) — ark read, CMVN, the 11x larger spliced matrix (processing/feature_reader.py:42-60)
and Trainer.update's pad-to-max stacking (neuralNetworks/trainer.py:276-307), all on the training thread.  Here the
host only concatenates the raw [T_u, D] matrices (CMVN and splicing run on the device, tfk_train_step_raw), and it
does so one or more batches AHEAD of the step in flight.

The dispenser's cursor semantics are kept: `return_batch` / `skip_batch` (validation rollback, resume;
nnet.py:104-105, 181-182) first un-read whatever was prefetched beyond the batch the trainer last consumed.
"""
import queue
import threading

import numpy as np
import torch


class RawBatch(object):
    """one packed batch: `utts` utterances / `frames` rows, cut into micro-batches of `utts_per_microbatch`"""

    __slots__ = ("slot", "frames", "utts", "feat_dim", "mb_rows", "raw", "labels", "offsets", "cmvn", "device_slot", "on_device",
                 "cursor_before")

    def microbatches(self):
        """(raw rows, labels, rebased utterance offsets, cmvn) views per micro-batch — device views once staged"""
        src = self.on_device if self.on_device is not None else (self.raw, self.labels, self.offsets, self.cmvn)
        raw, labels, offsets, cmvn = src
        n = offsets.shape[1] - 1
        for m in range(len(self.mb_rows) - 1):
            a, b = self.mb_rows[m], self.mb_rows[m + 1]
            yield raw[a:b], labels[a:b], offsets[m], cmvn[m * n:(m + 1) * n]


def cmvn_coefficients(stats):
    """(mean, 1/std) float32 [2, D] from accumulated statistics [2, D+1]: the formulas of apply_cmvn
    (processing/feature_reader.py:109-115) with the division turned into a multiplication for the device"""
    count = stats[0, -1]
    mean = stats[0, :-1] / count
    out = np.empty((2, stats.shape[1] - 1), np.float32)
    out[0] = mean
    out[1] = 1.0 / np.sqrt(stats[1, :-1] / count - np.square(mean))
    return out


class RawBatchFeeder(object):
    def __init__(self, dispenser, utts_per_microbatch=None, capacity_frames=None, depth=4, device=None):
        """dispenser: AlignmentBatchDispenser (its get_raw_batch is the source); utts_per_microbatch: the trainer's
        numutterances_per_minibatch (default: the whole batch); capacity_frames: rows a batch may hold (default
        batch size x max_input_length); depth: pinned slots (1 computing + 1 copying + ready/packing ones)"""
        self.dispenser = dispenser
        self.size = int(dispenser.size)
        self.n = int(utts_per_microbatch or self.size)
        if self.size % self.n != 0:
            raise ValueError("the batch size (%d) must be a multiple of numutterances_per_minibatch (%d)" % (self.size, self.n))
        self.context_width = int(dispenser.feature_reader.context_width)
        self._reader = dispenser.feature_reader.reader  # the ArkReader whose scp cursor the dispenser moves
        self.capacity = int(capacity_frames or self.size * dispenser.max_input_length)
        self.device = device
        self.depth = max(2, int(depth))
        self._slots = None
        self._free = queue.Queue()
        self._ready = queue.Queue()
        self._lock = threading.Lock()  # held by whoever moves the dispenser's cursor
        self._thread = None
        self._stop = False
        self._error = None
        self._coeff = {}
        self._staged = None  # a batch whose H2D was queued ahead of its use
        self._dev = None
        self.frames_out = 0  # frames handed to the trainer so far (bench.py)
        # feature dimension of the archive: ArkReader reads by index without moving the cursor
        self.feat_dim = int(self._reader.read_utt_data(0).shape[1])

    # ------------------------------------------------------------------ host side (no CUDA needed)
    def _allocate(self):
        pin, feat_dim = torch.cuda.is_available(), self.feat_dim
        mbs = self.size // self.n
        self._slots = [(torch.empty((self.capacity, feat_dim), dtype=torch.float32, pin_memory=pin),
                        torch.empty((self.capacity,), dtype=torch.int32, pin_memory=pin),
                        torch.empty((mbs, self.n + 1), dtype=torch.int32, pin_memory=pin),
                        torch.empty((self.size, 2, feat_dim), dtype=torch.float32, pin_memory=pin)) for _ in range(self.depth)]
        for k in range(self.depth):
            self._free.put(k)

    def _pack(self, slot, mats, stats, targets):
        raw_t, lab_t, off_t, cmvn_t = self._slots[slot]
        raw, lab, off, cmvn = raw_t.numpy(), lab_t.numpy(), off_t.numpy(), cmvn_t.numpy()
        frames = sum(m.shape[0] for m in mats)
        if frames > self.capacity:
            raise ValueError("batch of %d frames exceeds the feeder capacity %d" % (frames, self.capacity))
        np.concatenate(mats, axis=0, out=raw[:frames])
        row, mb_rows = 0, [0]
        for i, (m, st, t) in enumerate(zip(mats, stats, targets)):
            if t.shape[0] != m.shape[0]:
                raise ValueError("utterance %d of the batch: %d frames but %d targets" % (i, m.shape[0], t.shape[0]))
            key = id(st)
            hit = self._coeff.get(key)
            if hit is None or hit[0] is not st:
                hit = (st, cmvn_coefficients(st))  # the statistics object is kept so its id cannot be recycled
                self._coeff[key] = hit
            cmvn[i] = hit[1]
            lab[row:row + t.shape[0]] = t  # uint32 -> int32 (placeholder dtype, trainer.py:52-55)
            off[i // self.n, i % self.n] = row - mb_rows[i // self.n]
            row += m.shape[0]
            if (i + 1) % self.n == 0:
                off[i // self.n, self.n] = row - mb_rows[i // self.n]
                mb_rows.append(row)
        b = RawBatch()
        b.slot, b.frames, b.utts, b.feat_dim, b.mb_rows = slot, frames, len(mats), raw.shape[1], mb_rows
        b.raw, b.labels, b.offsets, b.cmvn = raw_t[:frames], lab_t[:frames], off_t, cmvn_t
        b.device_slot, b.on_device = None, None
        return b

    def _run(self):
        try:
            while True:
                slot = self._free.get()
                if slot is None or self._stop:
                    return
                with self._lock:
                    if self._stop:
                        return
                    cursor = self._reader.scp_position
                    mats, stats, targets = self.dispenser.get_raw_batch()
                    batch = self._pack(slot, mats, stats, targets)
                    batch.cursor_before = cursor  # where the scp cursor stood before this batch was read
                    self._ready.put(batch)
        except BaseException as exc:  # surfaced on the training thread by get()
            self._error = exc
            self._ready.put(None)

    def start(self):
        if self._thread is None:
            self._stop = False
            if self._slots is None:
                self._allocate()
            self._thread = threading.Thread(target=self._run, name="tfkaldi-feeder", daemon=True)
            self._thread.start()
        return self

    def get(self):
        """the next packed batch (host side); blocks until the background thread has one"""
        self.start()
        batch = self._ready.get()
        if batch is None:
            raise RuntimeError("feeder thread failed") from self._error
        return batch

    def release(self, batch):
        """the batch's buffers may be overwritten (its step has consumed them)"""
        if batch.device_slot is not None:
            self._dev["copied"][batch.device_slot].synchronize()  # the pinned slot must have left the host
        self._free.put(batch.slot)

    def _unread(self):
        """hand back everything that was prefetched but not consumed: the scp cursor goes back to exactly where it
        stood before the oldest unconsumed batch was read, i.e. where the synchronous loop would be now (the
        dispenser's own return_batch is NOT used for this: it carries the reference's off-by-one cursor quirks,
        processing/ark.py:144-149, which must only apply to the moves the trainer asks for).  Returns the number of
        batches un-read.  Call with the lock held."""
        pending = []
        if self._staged is not None:
            if self._staged.device_slot is not None:
                self._dev["copied"][self._staged.device_slot].synchronize()  # its copy may still be reading the pinned slot
            pending.append(self._staged)
            self._staged = None
        while True:
            try:
                batch = self._ready.get_nowait()
            except queue.Empty:
                break
            if batch is not None:
                pending.append(batch)
        if pending:
            self._reader.scp_position = pending[0].cursor_before
            for batch in pending:
                self._free.put(batch.slot)
        return len(pending)

    def return_batch(self):
        """dispenser.return_batch() as seen from the trainer: one batch back from the last CONSUMED one"""
        with self._lock:
            self._unread()
            self.dispenser.return_batch()

    def skip_batch(self):
        with self._lock:
            self._unread()
            self.dispenser.skip_batch()

    def get_batch(self):
        """the dispenser's own (spliced, host-side) batch at the trainer's position — e.g. for the validation set"""
        with self._lock:
            self._unread()
            return self.dispenser.get_batch()

    def split(self):
        with self._lock:
            self._unread()
            self.dispenser.split()

    def compute_target_count(self):
        return self.dispenser.compute_target_count()

    num_batches = property(lambda self: self.dispenser.num_batches)
    num_utt = property(lambda self: self.dispenser.num_utt)
    num_labels = property(lambda self: self.dispenser.num_labels)
    max_input_length = property(lambda self: self.dispenser.max_input_length)
    max_target_length = property(lambda self: self.dispenser.max_target_length)

    def close(self):
        """stop the thread and un-read what it had prefetched: the dispenser's cursor is left right after the last
        batch the trainer consumed, as if get_batch had been called synchronously all along"""
        self._stop = True
        self._free.put(None)
        if self._thread is not None:
            self._thread.join(timeout=30)
            self._thread = None
        with self._lock:
            self._unread()
        if self._slots is not None:  # no batch is outstanding any more: every slot is free again
            self._free = queue.Queue()
            for k in range(self.depth):
                self._free.put(k)

    # ------------------------------------------------------------------ device side
    def _device_buffers(self):
        if self._dev is None:
            dev = self.device if self.device is not None else torch.cuda.current_device()
            dev = torch.device("cuda", dev) if isinstance(dev, int) else torch.device(dev)
            mbs = self.size // self.n
            self._dev = {
                "device": dev,
                "bufs": [(torch.empty((self.capacity, self.feat_dim), dtype=torch.float32, device=dev),
                          torch.empty((self.capacity,), dtype=torch.int32, device=dev),
                          torch.empty((mbs, self.n + 1), dtype=torch.int32, device=dev),
                          torch.empty((self.size, 2, self.feat_dim), dtype=torch.float32, device=dev)) for _ in range(2)],
                "stream": torch.cuda.Stream(device=dev),
                "copied": [torch.cuda.Event() for _ in range(2)],
                "consumed": [None, None],
                "turn": 0,
            }
        return self._dev

    def _stage(self, batch):
        """queue the batch's host->device copy on the copy stream (does not wait for it)"""
        d = self._device_buffers()
        k = d["turn"]
        d["turn"] ^= 1
        raw, lab, off, cmvn = d["bufs"][k]
        with torch.cuda.stream(d["stream"]):
            if d["consumed"][k] is not None:
                d["stream"].wait_event(d["consumed"][k])  # the step that read this device buffer has finished
            raw[:batch.frames].copy_(batch.raw, non_blocking=True)
            lab[:batch.frames].copy_(batch.labels, non_blocking=True)
            off.copy_(batch.offsets, non_blocking=True)
            cmvn.copy_(batch.cmvn, non_blocking=True)
            d["copied"][k].record(d["stream"])
        batch.device_slot = k
        batch.on_device = (raw[:batch.frames], lab[:batch.frames], off, cmvn)
        return batch

    def get_on_device(self, device=None):
        """the next batch with its device views; the compute stream is made to wait for the copy.  Call
        `stage_next()` once the step's kernels are queued, so that the next batch's copy is set up while they run."""
        if device is not None and self._dev is None:
            self.device = device
        batch = self._staged if self._staged is not None else self._stage(self.get())
        self._staged = None
        self.frames_out += batch.frames
        d = self._dev
        torch.cuda.current_stream(d["device"]).wait_event(d["copied"][batch.device_slot])
        return batch

    def stage_next(self):
        """if the batch after the one in flight is already packed, queue its host->device copy now: it overlaps the
        step that was just launched.  (Only this thread takes batches out of the queue: no lock needed, the packer
        may hold it for a whole batch.)"""
        if self._staged is not None:
            return
        try:
            nxt = self._ready.get_nowait()
        except queue.Empty:
            return
        if nxt is None:
            raise RuntimeError("feeder thread failed") from self._error
        self._staged = self._stage(nxt)

    def consumed(self, batch):
        """call after the step's kernels have been enqueued: marks the device buffer and recycles the pinned slot"""
        d = self._dev
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(d["device"]))
        d["consumed"][batch.device_slot] = ev
        self.release(batch)
