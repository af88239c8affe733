"""One transport object, backend picked by registry key -- the role of zeus/common/ipc/uni_comm.py:27-80.

A thin proxy: the transport verbs are forwarded to the backend untouched (same positional / keyword calls as the
reference's callers make: ``send(data, name=None, block=True)``, ``recv(name=None, block=True)``, ``send_bytes``,
``recv_bytes(block=True)``, ``send_multipart``, ``recv_multipart``, ``delete``); ``info`` and ``close`` are the two members
with behaviour of their own."""
import logging
import threading

from ..registry import Registers

_VERBS = frozenset(("send", "recv", "send_bytes", "recv_bytes", "send_multipart", "recv_multipart", "delete"))


class UniComm(object):
    def __init__(self, comm_name, **comm_info):
        backend_cls = Registers.comm[comm_name]
        self.comm = backend_cls(comm_info)
        self.lock = threading.Lock()

    def __getattr__(self, verb):
        # only reached for names that are not real attributes: forward the transport verbs, refuse everything else
        if verb in _VERBS:
            return getattr(self.comm, verb)
        raise AttributeError(verb)

    @property
    def info(self):
        return str(self.comm)

    def close(self):
        with self.lock:
            closer = getattr(self.comm, "close", None)
            if closer is None:
                logging.info("comm backend %s has no close()", type(self.comm).__name__)
                return
            closer()
