"""zmq transport with the surface of zeus/common/ipc/comm_by_zmq.py:30-110; ``pyarrow.serialize`` (gone from current
pyarrow) is replaced by pickle protocol 5 with out-of-band numpy buffers sent as extra zmq frames (zero-copy send)."""
import pickle

import zmq

from ..registry import Registers

ZMQ_MIN_PORT = 20000
ZMQ_MAX_PORT = 40000


@Registers.comm
class CommByZmq(object):
    def __init__(self, comm_info):
        addr = comm_info.get("addr", "*")
        port = comm_info.get("port")
        zmq_type = comm_info.get("type", "PUB")
        comm_type = {"PUB": zmq.PUB, "SUB": zmq.SUB, "PUSH": zmq.PUSH, "PULL": zmq.PULL, "REP": zmq.REP, "REQ": zmq.REQ}.get(zmq_type)
        self._ctx = zmq.Context()
        socket = self._ctx.socket(comm_type)
        self._type = zmq_type
        self.bound_port = None
        if "*" in addr:
            self.bound_port = socket.bind_to_random_port("tcp://*", min_port=ZMQ_MIN_PORT, max_port=ZMQ_MAX_PORT, max_tries=100)
        else:
            socket.connect("tcp://" + str(addr) + ":" + str(port))
        if zmq_type == "SUB":
            socket.setsockopt(zmq.SUBSCRIBE, b"")
        self.socket = socket

    def send(self, ctr_info, data, name=None, block=True):
        bufs = []
        head = pickle.dumps((ctr_info, data), protocol=5, buffer_callback=bufs.append)
        self.socket.send_multipart([head] + [b.raw() for b in bufs], copy=False)

    def recv(self, name=None, block=True):
        msg = self.socket.recv_multipart(copy=False)
        ctr_info, data = pickle.loads(msg[0].buffer, buffers=[m.buffer for m in msg[1:]])
        return ctr_info, data

    def send_bytes(self, ctr_info, data):
        self.socket.send_multipart([ctr_info, data], copy=False)

    def recv_bytes(self, block=True):
        recv_data = self.socket.recv_multipart()
        return recv_data[0], recv_data[1]

    def __str__(self):
        return str({"port": self.bound_port, "type": self._type})

    def close(self):
        if self.socket:
            self.socket.close(linger=0)
            self.socket = None
        self._ctx.term()
