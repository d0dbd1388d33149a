"""Kaldi binary ark/scp IO with the interface of the reference's processing/ark.py.

Same classes, method names, return values and on-disk bytes as the reference (ArkReader
processing/ark.py:28-165, ArkWriter processing/ark.py:167-216), including its quirks:
  * entries are written as `<key>` immediately followed by `\\0BFM ` (no space after the key), so the
    archive is only consumable through the scp offsets, which point at the `\\0` (ark.py:204-210);
  * `split()` drops what was read so far AND the last utterance and leaves the cursor where it was
    (ark.py:161-165);
  * `read_previous_scp` wraps only once the cursor is already negative (ark.py:144-149).
What differs is how: the reader parses headers out of one shared memory map per archive instead of
re-opening the file for every utterance, and the writer keeps one buffered append handle per archive.
"""
import mmap
import os
import struct

import numpy as np

_HDR = struct.Struct("<xcccc")  # \0 B {F|D|C..} M ' '
_DIM = struct.Struct("<bi")  # \4 int32


class ArkReader(object):
    """Random access to binary float/double matrices through an .scp index."""

    def __init__(self, scp_path):
        self.scp_position = 0
        self.utt_ids = []
        self.scp_data = []
        with open(scp_path, "r") as scp:
            for line in scp:
                line = line.replace("\n", "")
                if line == "":
                    break  # the reference stops at the first empty line
                utt_id, where = line.split(" ")
                path, pos = where.split(":")
                self.utt_ids.append(utt_id)
                self.scp_data.append((path, pos))
        self._maps = {}

    def _view(self, path):
        entry = self._maps.get(path)
        size = os.path.getsize(path)
        if entry is None or entry[1] != size:
            with open(path, "rb") as fid:
                entry = (mmap.mmap(fid.fileno(), 0, access=mmap.ACCESS_READ), size)
            self._maps[path] = entry
        return entry[0]

    def read_utt_data(self, index):
        """matrix of utterance number `index` (float32 for 'F' archives, float64 for 'D')."""
        path, pos = self.scp_data[index]
        buf, off = self._view(path), int(pos)
        kind = _HDR.unpack_from(buf, off)
        if kind[0] != b"B":
            print("Input .ark file is not binary")
            exit(1)
        if kind[1] == b"C":
            print("Input .ark file is compressed")
            exit(1)
        _, rows = _DIM.unpack_from(buf, off + 5)
        _, cols = _DIM.unpack_from(buf, off + 10)
        dtype = np.float32 if kind[1] == b"F" else np.float64
        flat = np.frombuffer(buf, dtype=dtype, count=rows * cols, offset=off + 15)
        return flat.reshape(rows, cols)

    def read_next_utt(self):
        """(utt_id, matrix, looped): the next utterance, wrapping to the start at the end."""
        if len(self.scp_data) == 0:
            return None, None, True
        looped = self.scp_position >= len(self.scp_data)
        if looped:
            self.scp_position = 0
        self.scp_position += 1
        here = self.scp_position - 1
        return self.utt_ids[here], self.read_utt_data(here), looped

    def read_next_scp(self):
        """advance the cursor and return only the utterance id"""
        if self.scp_position >= len(self.scp_data):
            self.scp_position = 0
        self.scp_position += 1
        return self.utt_ids[self.scp_position - 1]

    def read_previous_scp(self):
        """move the cursor back and return the id it just left"""
        if self.scp_position < 0:
            self.scp_position = len(self.scp_data) - 1
        self.scp_position -= 1
        return self.utt_ids[self.scp_position + 1]

    def read_utt(self, utt_id):
        return self.read_utt_data(self.utt_ids.index(utt_id))

    def split(self):
        """split off the part that was read so far (and, as the reference does, the last entry)"""
        self.scp_data = self.scp_data[self.scp_position:-1]
        self.utt_ids = self.utt_ids[self.scp_position:-1]


class ArkWriter(object):
    """Append float32 matrices to a binary ark and index them in an scp."""

    def __init__(self, scp_path, default_ark):
        self.scp_path = scp_path
        self.scp_file_write = open(self.scp_path, "w")
        self.default_ark = default_ark
        self._arks = {}  # path -> [handle, next offset]

    def _ark(self, path):
        slot = self._arks.get(path)
        if slot is None:
            handle = open(path, "ab", buffering=1 << 22)  # append: the caller deletes stale archives
            slot = [handle, os.path.getsize(path)]
            self._arks[path] = slot
        return slot

    def write_next_utt(self, utt_id, utt_mat, ark_path=None):
        ark = ark_path or self.default_ark
        slot = self._ark(ark)
        handle = slot[0]
        utt_mat = np.ascontiguousarray(utt_mat, dtype=np.float32)
        rows, cols = utt_mat.shape
        key = utt_id.encode()
        pos = slot[1] + len(key)  # scp offset: the \0 that starts the binary marker
        handle.write(key + _HDR.pack(b"B", b"F", b"M", b" ") + _DIM.pack(4, rows) + _DIM.pack(4, cols))
        handle.write(memoryview(utt_mat).cast("B"))
        slot[1] = pos + 15 + utt_mat.nbytes
        # the reference reopened and closed the archive for every utterance (ark.py:201,211), so an entry was readable
        # as soon as its scp line existed: keep that — archive bytes reach the file before the line that points at them
        handle.flush()
        self.scp_file_write.write("%s %s:%s\n" % (utt_id, ark, pos))
        self.scp_file_write.flush()

    # ---- streaming variant of write_next_utt (same bytes): the caller fills the matrix in pieces, from any thread
    def begin_utt(self, utt_id, rows, cols, ark_path=None):
        """Write the key and the matrix header of a rows x cols float32 entry and reserve the space of its data.
        Returns a handle for write_rows / finish_utt.  Used by the decoder's output pipeline: the log-likelihoods of a
        long utterance arrive tile by tile from the GPU (774 MB for 100 000 frames x 1936 pdf-ids) and are written with
        positional writes straight out of pinned memory while later tiles are still being computed."""
        ark = ark_path or self.default_ark
        slot = self._ark(ark)
        handle = slot[0]
        key = utt_id.encode()
        pos = slot[1] + len(key)
        handle.write(key + _HDR.pack(b"B", b"F", b"M", b" ") + _DIM.pack(4, rows) + _DIM.pack(4, cols))
        handle.flush()
        data_offset = pos + 15
        nbytes = int(rows) * int(cols) * 4
        slot[1] = data_offset + nbytes
        if len(slot) < 3:  # a second descriptor for mapping (the append handle cannot be mapped for writing)
            slot.append(os.open(ark, os.O_RDWR))
        os.ftruncate(slot[2], slot[1])  # the append handle continues after the reserved region
        # The reserved region is filled through a shared mapping: write()/pwrite() on ONE file serialise on the
        # inode lock whatever the number of threads (~2-4 GB/s on tmpfs), page-granular copies into a mapping do not.
        entry = {"utt_id": utt_id, "ark": ark, "pos": pos, "data_offset": data_offset, "rows": int(rows), "cols": int(cols),
                 "map": None, "bytes": None}
        if nbytes:
            base = data_offset - data_offset % mmap.ALLOCATIONGRANULARITY
            entry["map"] = mmap.mmap(slot[2], data_offset + nbytes - base, offset=base)
            entry["bytes"] = np.frombuffer(entry["map"], dtype=np.uint8, offset=data_offset - base)
        return entry

    @staticmethod
    def write_rows(entry, first_row, block):
        """rows [first_row, first_row + len(block)) of a begin_utt entry; block: C-contiguous float32 [n, cols].
        Thread-safe (the copy releases the GIL): several blocks of one entry may be written concurrently."""
        block = np.ascontiguousarray(block, dtype=np.float32)
        lo = int(first_row) * entry["cols"] * 4
        np.copyto(entry["bytes"][lo:lo + block.nbytes], block.reshape(-1).view(np.uint8))

    def finish_utt(self, entry):
        """all rows are written: index the entry (scp lines appear in finish order)"""
        if entry["map"] is not None:
            entry["bytes"] = None  # drop the exported buffer before closing the mapping
            entry["map"].close()
            entry["map"] = None
        self.scp_file_write.write("%s %s:%s\n" % (entry["utt_id"], entry["ark"], entry["pos"]))
        self.scp_file_write.flush()

    def flush(self):
        for slot in self._arks.values():
            slot[0].flush()
        self.scp_file_write.flush()

    def close(self):
        for slot in self._arks.values():
            slot[0].close()
            if len(slot) > 2:
                os.close(slot[2])
        self._arks = {}
        self.scp_file_write.close()
