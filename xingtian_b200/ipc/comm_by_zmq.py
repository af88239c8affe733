"""zmq transport behind ``Registers.comm`` (role of zeus/common/ipc/comm_by_zmq.py:30-110).

Same construction contract -- ``comm_info`` = {type: PUB/SUB/PUSH/PULL/REP/REQ, addr, port}; an address containing ``*``
binds a random port in [20000, 40000), anything else connects -- and the same four methods.  The wire format differs:
``pyarrow.serialize`` no longer exists, so a message is one pickle-5 header frame followed by the numpy buffers as raw
zmq frames (sent without copying)."""
import pickle

import zmq

from ..registry import Registers

PORT_RANGE = (20000, 40000)
_SOCKET_KIND = dict(PUB=zmq.PUB, SUB=zmq.SUB, PUSH=zmq.PUSH, PULL=zmq.PULL, REP=zmq.REP, REQ=zmq.REQ)


@Registers.comm
class CommByZmq(object):
    def __init__(self, comm_info):
        self.kind = comm_info.get("type", "PUB")
        if self.kind not in _SOCKET_KIND:
            raise KeyError("unknown zmq socket type: {}".format(self.kind))
        self._ctx = zmq.Context()
        self.socket = self._ctx.socket(_SOCKET_KIND[self.kind])
        self.port = self._attach(str(comm_info.get("addr", "*")), comm_info.get("port"))
        if self.kind == "SUB":
            self.socket.setsockopt(zmq.SUBSCRIBE, b"")

    def _attach(self, addr, port):
        """Server side (wildcard address): bind a free port and report it; client side: connect to addr:port."""
        if "*" in addr:
            return self.socket.bind_to_random_port("tcp://*", min_port=PORT_RANGE[0], max_port=PORT_RANGE[1], max_tries=100)
        self.socket.connect("tcp://{}:{}".format(addr, port))
        return None

    # -- python objects ----------------------------------------------------------------------------------------------
    def send(self, ctr_info, data, name=None, block=True):
        frames = []
        header = pickle.dumps((ctr_info, data), protocol=5, buffer_callback=frames.append)
        self.socket.send_multipart([header] + [f.raw() for f in frames], copy=False)

    def recv(self, name=None, block=True):
        parts = self.socket.recv_multipart(copy=False)
        return pickle.loads(parts[0].buffer, buffers=[p.buffer for p in parts[1:]])       # (ctr_info, data)

    # -- raw buffers -------------------------------------------------------------------------------------------------
    def send_bytes(self, ctr_info, data):
        self.socket.send_multipart([ctr_info, data], copy=False)

    def recv_bytes(self, block=True):
        head, body = self.socket.recv_multipart()[:2]
        return head, body

    def __str__(self):
        return str({"port": self.port, "type": self.kind})

    def close(self):
        sock, self.socket = self.socket, None
        if sock is not None:
            sock.close(linger=0)
            self._ctx.term()
