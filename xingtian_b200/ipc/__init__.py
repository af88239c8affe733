"""Comm backends behind the reference's ``Registers.comm`` / ``UniComm`` surface."""
from .share_by_shm import ShareByShm, ShareByPlasma     # noqa: F401
from .comm_by_zmq import CommByZmq                      # noqa: F401
from .uni_comm import UniComm                           # noqa: F401
