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


def _explorer(req, rep, explorer_id, frames, out_q):
    from xingtian_b200.service.predictor import RemotePredict
    stub = RemotePredict(req, rep, explorer_id=explorer_id, broker_id=0)
    out_q.put((explorer_id, [float(stub.predict(f)) for f in frames]))


def test_remote_predict_from_explorer_processes_over_shared_memory():
    """f2 + f3 together on CPU: two explorer PROCESSES hold no model, send frames through the shared-memory store with
    RemotePredict and block on their own reply channel; the predictor (stub model) answers batched; replies are routed by
    explorer id the way the broker does it (xt/framework/broker.py:401-419)."""
    from xingtian_b200.ipc import UniComm
    from xingtian_b200.service.predictor import BatchedPredictor

    class StubAlg(object):
        class actor(object):
            state_dim = (84, 84, 4)

    sizes = []

    def predict_fn(batch):
        sizes.append(batch.shape[0])
        return batch.reshape(batch.shape[0], -1).astype(np.float64).sum(1)

    class Router(object):                      # reply side of the broker: one channel per explorer
        def __init__(self, chans):
            self.chans = chans

        def send(self, msg):
            self.chans[msg["ctr_info"]["explorer_id"]].send(msg)

    req = UniComm("ShareByPlasma")
    reps = {i: UniComm("ShareByPlasma") for i in (0, 1)}
    pred = BatchedPredictor(0, StubAlg(), req, Router(reps), predict_fn=predict_fn, max_batch=16, max_wait_s=0.02, report_period=10 ** 9)
    rng = np.random.default_rng(4)
    frames = {i: [rng.integers(0, 256, (84, 84, 4), dtype=np.uint8) for _ in range(3)] for i in (0, 1)}
    ctx = mp.get_context("fork")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_explorer, args=(req, reps[i], i, frames[i], out_q)) for i in (0, 1)]
    for p in procs:
        p.start()
    answered = 0
    import time
    deadline = time.time() + 30
    while answered < 6 and time.time() < deadline:
        answered += pred.process_once(timeout=0.5)
    got = dict(out_q.get(timeout=10) for _ in range(2))
    for p in procs:
        p.join(10)
        assert p.exitcode == 0
    assert answered == 6 and sum(sizes) == 6 and max(sizes) <= 2
    for i in (0, 1):
        assert got[i] == [float(f.astype(np.float64).sum()) for f in frames[i]]
    req.close()
    for c in reps.values():
        c.close()
