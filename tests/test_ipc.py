"""f3: the comm backends behind the reference's Registers.comm / UniComm surface -- shared-memory object store
(ShareByPlasma key) across processes, zmq PUSH/PULL, and a BatchedPredictor served over the store."""
import multiprocessing as mp

import numpy as np


def _producer(comm, n):
    rng = np.random.default_rng(0)
    for i in range(n):
        frames = rng.integers(0, 256, (16, 84, 84, 4), dtype=np.uint8)
        msg = {"data": {"cur_state": frames, "action": np.arange(16, dtype=np.int32) + i, "info": [{"k": i}]},
               "ctr_info": {"cmd": "train", "broker_id": 1, "explorer_id": i, "agent_id": 0}}
        comm.send(msg)
        assert msg["data"] == {}          # sender's copy is dropped like share_by_plasma.py:66-75


def test_share_by_shm_train_messages_across_processes():
    from xingtian_b200.ipc import UniComm
    from xingtian_b200.registry import Registers
    assert "ShareByPlasma" in Registers.comm and "CommByZmq" in Registers.comm
    comm = UniComm("ShareByPlasma", size=1 << 20)
    p = mp.get_context("fork").Process(target=_producer, args=(comm, 3))
    p.start()
    rng = np.random.default_rng(0)
    for i in range(3):
        ctr, data = comm.recv()
        assert ctr["cmd"] == "train" and ctr["explorer_id"] == i and "object_id" in ctr
        np.testing.assert_array_equal(data["cur_state"], rng.integers(0, 256, (16, 84, 84, 4), dtype=np.uint8))
        np.testing.assert_array_equal(data["action"], np.arange(16, dtype=np.int32) + i)
        assert data["info"] == [{"k": i}] and data["cur_state"].flags.writeable
    p.join(10)
    assert p.exitcode == 0
    assert comm.recv(block=False) is None
    # raw-buffer and multipart forms
    comm.send_bytes(b"abc" * 1000, "model")
    ctr, buf = comm.recv_bytes(True)
    assert bytes(buf) == b"abc" * 1000 and ctr["data_type"] == "model"
    comm.delete(ctr["object_id"])
    comm.send_multipart([b"x" * 10, b"y" * 20])
    assert [bytes(b) for b in comm.recv_multipart()] == [b"x" * 10, b"y" * 20]
    comm.close()


def test_comm_by_zmq_push_pull_roundtrip():
    from xingtian_b200.ipc import UniComm
    pull = UniComm("CommByZmq", type="PULL")
    port = eval(pull.info)["port"]
    push = UniComm("CommByZmq", type="PUSH", addr="127.0.0.1", port=port)
    w = {"pi_latent/kernel": np.random.default_rng(1).standard_normal((256, 4)).astype(np.float32)}
    push.comm.send({"cmd": "predict", "sub_cmd": "sync_weights"}, w)
    ctr, data = pull.recv()
    assert ctr["sub_cmd"] == "sync_weights"
    np.testing.assert_array_equal(data["pi_latent/kernel"], w["pi_latent/kernel"])
    push.send_bytes(b"hdr", b"payload")
    assert tuple(bytes(x) for x in pull.recv_bytes()) == (b"hdr", b"payload")
    push.close(); pull.close()


def test_batched_predictor_over_the_shared_memory_store():
    """Requests of several explorers travel through the store, are answered by ONE batched call, replies are routed."""
    from xingtian_b200.ipc import UniComm
    from xingtian_b200.service.predictor import BatchedPredictor, make_msg

    class StubAlg(object):
        class actor(object):
            state_dim = (84, 84, 4)
        calls = []

        def set_weights(self, w):
            self.w = w

    alg = StubAlg()

    def predict_fn(batch):
        alg.calls.append(batch.shape[0])
        return batch.reshape(batch.shape[0], -1).sum(1).astype(np.float64)

    req, rep = UniComm("ShareByShm"), UniComm("ShareByShm")
    pred = BatchedPredictor(0, alg, req, rep, predict_fn=predict_fn, max_batch=8, max_wait_s=0.05)
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (5, 84, 84, 4), dtype=np.uint8)
    for i in range(5):
        req.send(make_msg(frames[i].copy(), cmd="predict", sub_cmd="predict", explorer_id=i, broker_id=7))
    answered = pred.process_once(timeout=1.0)
    assert answered == 5 and alg.calls == [5]
    got = {}
    for _ in range(5):
        ctr, data = rep.recv()
        got[ctr["explorer_id"]] = data
        assert ctr["cmd"] == "predict_reply" and ctr["broker_id"] == 7
    for i in range(5):
        assert got[i] == float(frames[i].sum())
    req.close(); rep.close()
