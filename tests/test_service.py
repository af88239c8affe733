"""CPU tier: the batched predictor service against a stub model - request/reply contract of
xt/framework/predictor.py:47-87, batching, routing, weight-sync ordering, statistics message."""
import queue
import threading

import numpy as np

from xingtian_b200.service import BatchedPredictor, RemotePredict, make_msg


class _Actor(object):
    state_dim = [6, 5, 2]

    def __init__(self):
        self.bias = 0
        self.calls = []

    def predict(self, batch):          # PPO-shaped result: (action [B] int32, logp [B,1], value [B,1])
        batch = np.asarray(batch)
        assert batch.shape[1:] == (6, 5, 2)
        self.calls.append(batch.shape[0])
        s = batch.reshape(batch.shape[0], -1).sum(1).astype(np.int64) + self.bias
        return (s % 7).astype(np.int32), (-s[:, None]).astype(np.float32), (2.0 * s[:, None]).astype(np.float32)


class _Alg(object):
    def __init__(self):
        self.actor = _Actor()

    def set_weights(self, w):
        self.actor.bias = int(w["bias"])


def _expected(state, bias=0):
    s = int(np.asarray(state).sum()) + bias
    return s % 7, -float(s), 2.0 * s


def test_requests_are_batched_and_routed_back():
    alg, req = _Alg(), queue.Queue()
    replies = [queue.Queue() for _ in range(5)]

    class Router(object):               # what the broker does with predict_reply messages: route by explorer_id
        def put(self, msg):
            assert msg["ctr_info"]["cmd"] == "predict_reply" and msg["ctr_info"]["broker_id"] == 3
            replies[msg["ctr_info"]["explorer_id"]].put(msg)

    svc = BatchedPredictor(0, alg, req, Router(), max_wait_s=0.05)
    rng = np.random.default_rng(0)
    states = [rng.integers(0, 9, (6, 5, 2)) for _ in range(5)]
    for e, s in enumerate(states):      # all five are waiting before the window opens
        req.put(make_msg(s, cmd="predict", sub_cmd="predict", explorer_id=e, broker_id=3))
    assert svc.process_once(timeout=1.0) == 5
    assert alg.actor.calls == [5]       # ONE batched forward
    for e, s in enumerate(states):
        a, logp, v = replies[e].get_nowait()["data"]
        ea, el, ev = _expected(s)
        assert int(a) == ea and logp.shape == (1,) and float(logp[0]) == el and float(v[0]) == ev
    assert svc.process_once(timeout=0.01) == 0      # nothing waiting: times out


def test_vector_env_requests_and_max_batch():
    alg, req, rep = _Alg(), queue.Queue(), queue.Queue()
    svc = BatchedPredictor(0, alg, req, rep, max_batch=6, max_wait_s=0.05)
    rng = np.random.default_rng(1)
    big = [rng.integers(0, 9, (4, 6, 5, 2)) for _ in range(3)]          # three explorers with 4 envs each
    for e, s in enumerate(big):
        req.put(make_msg(s, cmd="predict", explorer_id=e))              # sub_cmd defaults to predict (predictor.py:54)
    assert svc.process_once(1.0) == 1      # 4 + 4 > 6: the second request waits for the next window
    assert svc.process_once(1.0) == 1
    assert svc.process_once(1.0) == 1
    assert alg.actor.calls == [4, 4, 4]
    for e, s in enumerate(big):
        msg = rep.get_nowait()
        assert msg["ctr_info"]["explorer_id"] == e
        a, logp, v = msg["data"]
        assert a.shape == (4,) and logp.shape == (4, 1)
        for i in range(4):
            assert int(a[i]) == _expected(s[i])[0]


def test_sync_weights_is_a_barrier_in_arrival_order():
    alg, req, rep = _Alg(), queue.Queue(), queue.Queue()
    svc = BatchedPredictor(0, alg, req, rep, max_wait_s=0.05)
    s0, s1 = np.ones((6, 5, 2), np.int64), np.full((6, 5, 2), 2, np.int64)
    req.put(make_msg(s0, cmd="predict", sub_cmd="predict", explorer_id=0))
    req.put((dict(cmd="predict", sub_cmd="sync_weights"), {"bias": 3}))          # UniComm (ctr_info, data) form
    req.put(make_msg(s1, cmd="predict", sub_cmd="predict", explorer_id=1))
    assert svc.process_once(1.0) == 1 and alg.actor.bias == 0     # answered with the old weights
    assert svc.process_once(1.0) == 0 and alg.actor.bias == 3     # the sync itself
    assert svc.process_once(1.0) == 1
    assert int(rep.get_nowait()["data"][0]) == _expected(s0, 0)[0]
    assert int(rep.get_nowait()["data"][0]) == _expected(s1, 3)[0]
    try:
        req.put(make_msg(s0, cmd="nonsense", sub_cmd="explode"))
        svc.process_once(1.0)
        raise AssertionError("an unknown sub-command must raise (predictor.py:62-66)")
    except KeyError:
        pass


def test_explorer_threads_through_the_stub_and_stats_message():
    alg, req = _Alg(), queue.Queue()
    n_exp, n_step = 8, 40
    inboxes = [queue.Queue() for _ in range(n_exp)]
    stats = []

    class Router(object):
        def put(self, msg):
            if msg["ctr_info"]["cmd"].startswith("stats_msg"):
                stats.append(msg)
            else:
                inboxes[msg["ctr_info"]["explorer_id"]].put(msg)

    svc = BatchedPredictor(0, alg, req, Router(), predictor_name="_p0", max_wait_s=0.002, report_period=100)
    stop = threading.Event()

    def serve():
        while not stop.is_set():
            svc.process_once(timeout=0.05)

    errors = []

    def explore(e):
        stub = RemotePredict(req, inboxes[e], explorer_id=e)
        rng = np.random.default_rng(100 + e)
        try:
            for _ in range(n_step):
                s = rng.integers(0, 9, (6, 5, 2))
                a, logp, v = stub.predict(s)
                assert int(a) == _expected(s)[0] and float(v[0]) == _expected(s)[2]
        except Exception as exc:      # noqa: BLE001
            errors.append(exc)

    t = threading.Thread(target=serve)
    t.start()
    workers = [threading.Thread(target=explore, args=(e,)) for e in range(n_exp)]
    [w.start() for w in workers]
    [w.join(30) for w in workers]
    stop.set()
    t.join(5)
    assert not errors, errors
    assert sum(alg.actor.calls) == n_exp * n_step
    assert max(alg.actor.calls) > 1                  # concurrent explorers were actually stacked
    assert stats and stats[0]["ctr_info"]["cmd"] == "stats_msg_p0"
    assert set(stats[0]["data"]) == {"mean_predictor_wait_ms", "mean_predictor_infer_ms", "mean_predictor_batch"}
