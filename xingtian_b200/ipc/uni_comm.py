"""UniComm facade (zeus/common/ipc/uni_comm.py:27-80): one object, backend chosen by registry key."""
import logging
import threading

from ..registry import Registers


class UniComm(object):
    def __init__(self, comm_name, **comm_info):
        self.comm = Registers.comm[comm_name](comm_info)
        self.lock = threading.Lock()

    def send(self, data, name=None, block=True, **kwargs):
        return self.comm.send(data, name, block, **kwargs)

    def recv(self, name=None, block=True):
        return self.comm.recv(name, block)

    def send_bytes(self, ctr_info, data):
        return self.comm.send_bytes(ctr_info, data)

    def recv_bytes(self, block=True):
        return self.comm.recv_bytes(block)

    def send_multipart(self, data):
        return self.comm.send_multipart(data)

    def recv_multipart(self):
        return self.comm.recv_multipart()

    def delete(self, name):
        return self.comm.delete(name)

    @property
    def info(self):
        return str(self.comm)

    def close(self):
        logging.debug("start close comm...")
        with self.lock:
            try:
                self.comm.close()
            except AttributeError as err:
                logging.info("call comm.close failed! with: \n%s", err)
