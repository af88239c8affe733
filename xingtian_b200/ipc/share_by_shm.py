"""Shared-memory object store with the method surface of the reference's plasma backend
(zeus/common/ipc/share_by_plasma.py:33-165).

The reference serialises with ``pyarrow.serialize`` into a ``plasma_store`` server process and lz4-compresses large
payloads; neither exists in current pyarrow.  Here an object is one POSIX shared-memory segment
(``multiprocessing.shared_memory``): the message is pickled with protocol 5 and its numpy buffers travel out of band,
i.e. uint8 frame stacks are copied once into the segment and once out, uncompressed (the learner wants the raw frames:
they go straight into the pinned staging ring).  Control information travels over a ``multiprocessing.Queue`` exactly
like the reference's ``control_q``.  Object ids are segment names.
"""
import os
import pickle
import struct
import sys
from multiprocessing import Queue, shared_memory

from ..registry import Registers

_HDR = struct.Struct("<QI")        # pickle length, number of out-of-band buffers


def _dumps(obj):
    bufs = []
    head = pickle.dumps(obj, protocol=5, buffer_callback=bufs.append)
    raws = [b.raw() for b in bufs]
    sizes = [r.nbytes for r in raws]
    total = _HDR.size + 8 * len(sizes) + len(head) + sum(sizes)
    return head, raws, sizes, total


def _untrack(seg):
    """multiprocessing's resource tracker unlinks every segment a process created (or, in 3.12, merely attached) when that
    process exits -- here the creator is an explorer that may exit while its message is still waiting in the store, and
    the segment's lifetime is managed explicitly (receiver unlinks / delete()).  Take the segment off the tracker's list."""
    try:
        from multiprocessing import resource_tracker
        resource_tracker.unregister(seg._name, "shared_memory")
    except Exception:          # tracker layout differs (3.13+: created with track=False instead)
        pass


def _put(obj):
    """Serialise `obj` into a fresh segment; returns its name (the object id)."""
    head, raws, sizes, total = _dumps(obj)
    seg = shared_memory.SharedMemory(create=True, size=max(total, 1))
    _untrack(seg)
    mv = seg.buf
    _HDR.pack_into(mv, 0, len(head), len(sizes))
    off = _HDR.size
    for n in sizes:
        struct.pack_into("<Q", mv, off, n); off += 8
    mv[off:off + len(head)] = head; off += len(head)
    for r, n in zip(raws, sizes):
        mv[off:off + n] = r.cast("B") if r.format != "B" or r.ndim != 1 else r; off += n
    name = seg.name
    seg.close()
    return name


def _get(name, unlink=True):
    seg = shared_memory.SharedMemory(name=name)         # 3.12 registers an attached segment too ...
    if not unlink:
        _untrack(seg)                                   # ... unlink() below takes it off again, otherwise do it here
    try:
        mv = seg.buf
        hlen, nb = _HDR.unpack_from(mv, 0)
        off = _HDR.size
        sizes = [struct.unpack_from("<Q", mv, off + 8 * i)[0] for i in range(nb)]
        off += 8 * nb
        head = bytes(mv[off:off + hlen]); off += hlen
        bufs = []
        for n in sizes:
            bufs.append(bytearray(mv[off:off + n])); off += n      # private copy: the segment goes away
        obj = pickle.loads(head, buffers=bufs)
        del mv
    finally:
        seg.close()
        if unlink:
            try:
                seg.unlink()
            except FileNotFoundError:
                pass
    return obj


class ShareByShm(object):
    """send / recv / send_bytes / recv_bytes / send_multipart / recv_multipart / delete / close like ShareByPlasma."""

    def __init__(self, comm_info):
        self.size_shared_mem = comm_info.get("size", 1000000000)       # accepted for config compatibility
        self.path = comm_info.get("path", "/xtb_shm" + str(os.getpid()))
        self.compress = False                                           # raw frames, no lz4
        self.control_q = Queue()
        self.client = {}
        self.start()

    # -- whole messages -------------------------------------------------------------------------------------------
    def send(self, data, name=None, block=True):
        """share_by_plasma.py:49-75: payload into the store, ctr_info (+ object id) onto the control queue, then the
        sender's copy of a train / predict payload is dropped."""
        object_id = _put(data["data"])
        ctr = dict(data["ctr_info"])
        ctr["object_id"] = object_id
        self.control_q.put(ctr)
        cmd_type = str(data["ctr_info"].get("cmd"))
        if cmd_type.startswith("train") and isinstance(data["data"], dict):
            for key in list(data["data"].keys()):
                del data["data"][key]
        elif cmd_type.startswith("predict"):
            del data["data"]

    def recv(self, name=None, block=True):
        """share_by_plasma.py:77-95 -> (ctr_info, data); None when non-blocking and nothing waits."""
        if not block and self.control_q.empty():
            return None
        ctr_info = self.control_q.get()
        data = _get(ctr_info["object_id"])
        return ctr_info, data

    # -- raw buffers ------------------------------------------------------------------------------------------------
    def send_bytes(self, data_buffer, data_type="data"):
        """share_by_plasma.py:97-101."""
        object_id = _put(bytes(data_buffer))
        self.control_q.put({"object_id": object_id, "data_type": data_type})

    def recv_bytes(self, block=True):
        """share_by_plasma.py:103-114: the object stays in the store until delete()."""
        if not block and self.control_q.empty():
            return None, None
        ctr_info = self.control_q.get()
        return ctr_info, _get(ctr_info["object_id"], unlink=False)

    def delete(self, object_id):
        """share_by_plasma.py:116-120."""
        try:
            seg = shared_memory.SharedMemory(name=object_id)
            seg.close(); seg.unlink()
        except FileNotFoundError:
            pass

    def send_multipart(self, data_buffer):
        """share_by_plasma.py:122-128."""
        self.control_q.put(len(data_buffer))
        for buf in data_buffer:
            self.control_q.put(_put(bytes(buf)))

    def recv_multipart(self):
        """share_by_plasma.py:130-142."""
        n = self.control_q.get()
        return [_get(self.control_q.get()) for _ in range(n)]

    # -- lifecycle (no server process to start) ----------------------------------------------------------------------
    def start(self):
        return None

    def connect(self):
        return self

    def close(self):
        while not self.control_q.empty():       # drop what was never received
            item = self.control_q.get()
            oid = item.get("object_id") if isinstance(item, dict) else (item if isinstance(item, str) else None)
            if oid:
                self.delete(oid)

    def __str__(self):
        return str({"type": "shm", "pid": os.getpid()})


@Registers.comm
class ShareByPlasma(ShareByShm):
    """Registered under the reference's key so ``UniComm("ShareByPlasma", ...)`` (zeus/common/ipc/uni_comm.py:30-33,
    xt/framework/broker.py:97-119) resolves to the shared-memory store."""


Registers.comm(ShareByShm)
assert sys.version_info >= (3, 8)
