"""Batched predictor service: many explorers, one batched forward.

The reference's ``Predictor`` (xt/framework/predictor.py:31-98) is a process that answers one request at a time:
``ctr_info, data = request_q.recv()`` -> ``alg.predict(state)`` on a CPU model at batch 1 -> ``reply_q.put(message(action,
cmd="predict_reply", broker_id, explorer_id))``, with a ``sync_weights`` sub-command and a stats message every 200
requests.  The explorer side (xt/agent/dqn/cartpole_dqn.py:51-55) sends ``message(state, cmd="predict")`` and blocks on
the reply.

Here the same request / reply contract feeds ONE batched call: requests that are waiting (up to ``max_batch`` samples,
or until ``max_wait_s`` has passed since the first one) are stacked along axis 0, evaluated by ``predict_fn`` (default
``alg.actor.predict``, i.e. ``xtb_ppo_predict_host``: one staged H2D, one CUDA graph, one packed D2H for the whole
batch) and the result rows are scattered back to their senders.  ``sync_weights`` is a barrier in arrival order:
requests received before it are answered with the old weights.
"""
import queue
import time

import numpy as np

__all__ = ["BatchedPredictor", "RemotePredict", "make_msg"]


def make_msg(data, **ctr):
    """Message layout of zeus/common/ipc/message.py:24-29: {"data": ..., "ctr_info": {broker_id, explorer_id, agent_id, cmd}}."""
    info = dict(broker_id=-1, explorer_id=-1, agent_id=-1, cmd="train")
    info.update(ctr)
    return {"data": data, "ctr_info": info}


class _Inbox(object):
    """Uniform (msg | None) = take(timeout) over queue.Queue / multiprocessing.Queue (``get``), multiprocessing
    connections (``poll`` + ``recv``) and UniComm-like objects that only offer a blocking ``recv``."""

    def __init__(self, q):
        self.q = q

    def take(self, timeout):
        q = self.q
        if hasattr(q, "get"):
            try:
                return q.get(True, timeout) if timeout is None or timeout > 0 else q.get(False)
            except queue.Empty:
                return None
        if hasattr(q, "poll"):
            if timeout is None or q.poll(timeout):
                return q.recv()
            return None
        if timeout is None:
            return q.recv()
        # UniComm-like transport (zeus/common/ipc/uni_comm.py:37-41): poll its non-blocking recv up to the timeout
        end = time.time() + timeout
        while True:
            try:
                got = q.recv(block=False)
            except TypeError:
                return None                                  # blocking-only transport: no opportunistic batching
            if got is not None or time.time() >= end:
                return got
            time.sleep(0.0002)


def _unpack(item):
    """UniComm.recv() yields (ctr_info, data) (predictor.py:50-51); dict messages are accepted as well."""
    if isinstance(item, dict):
        return item["ctr_info"], item["data"]
    ctr, data = item
    return ctr, data


def _rows(result, lo, hi, squeeze):
    """Rows [lo, hi) of a batched result (ndarray, or tuple/list of ndarrays sharing axis 0)."""
    if isinstance(result, (tuple, list)):
        return type(result)(_rows(r, lo, hi, squeeze) for r in result)
    part = np.asarray(result)[lo:hi]
    return part[0] if squeeze else part


class BatchedPredictor(object):
    def __init__(self, predictor_id, alg, request_q, reply_q, predictor_name="", predict_fn=None, max_batch=None,
                 max_wait_s=0.0005, report_period=200):
        self.predictor_id = predictor_id
        self.predictor_name = predictor_name
        self.alg = alg
        self.inbox = _Inbox(request_q)
        self.reply_q = reply_q
        self.predict_fn = predict_fn if predict_fn is not None else alg.actor.predict
        self.max_batch = int(max_batch) if max_batch else None
        self.max_wait_s = float(max_wait_s)
        self._report_period = int(report_period)
        self._held = None                  # a message taken from the inbox that must wait for the next round
        self.reset_stats()

    # -- statistics (zeus/common/util/profile_stats.py:98-124 plus the batch size) ----------------------------------
    def reset_stats(self):
        self.obs_wait_time = 0.0
        self.inference_time = 0.0
        self.iters = 0
        self.batches = 0

    def stats(self):
        n = max(self.iters, 1)
        out = {"mean_predictor_wait_ms": self.obs_wait_time * 1000.0 / n,
               "mean_predictor_infer_ms": self.inference_time * 1000.0 / n,
               "mean_predictor_batch": self.iters / float(max(self.batches, 1))}
        self.reset_stats()
        return out

    # -- one batching window ------------------------------------------------------------------------------------------
    def _send(self, msg):
        (self.reply_q.put if hasattr(self.reply_q, "put") else self.reply_q.send)(msg)

    def process_once(self, timeout=None):
        """Serve one window: wait (``timeout``; None = forever) for a first message, gather what else is already
        waiting, answer.  Returns the number of predict requests answered (0: timed out, or a weight sync)."""
        t0 = time.time()
        first = self._held if self._held is not None else self.inbox.take(timeout)
        self._held = None
        if first is None:
            return 0
        self.obs_wait_time += time.time() - t0
        ctr, data = _unpack(first)
        cmd = ctr.get("sub_cmd", "predict")
        if cmd == "sync_weights":
            self.alg.set_weights(data)
            return 0
        if cmd != "predict":
            raise KeyError("invalid cmd: {}".format(ctr.get("cmd")))
        pending = [(ctr, np.asarray(data))]
        item_shape = tuple(getattr(self.alg.actor, "state_dim", np.asarray(data).shape))
        count = lambda a: 1 if a.shape == item_shape else a.shape[0]
        total = count(pending[0][1])
        deadline = time.time() + self.max_wait_s
        while self.max_batch is None or total < self.max_batch:
            nxt = self.inbox.take(max(0.0, deadline - time.time()))
            if nxt is None:
                break
            c2, d2 = _unpack(nxt)
            if c2.get("sub_cmd", "predict") != "predict":
                self._held = nxt            # barrier: answer the gathered requests first
                break
            d2 = np.asarray(d2)
            if self.max_batch is not None and total + count(d2) > self.max_batch:
                self._held = nxt
                break
            pending.append((c2, d2))
            total += count(d2)
        t1 = time.time()
        batch = np.concatenate([d if d.shape != item_shape else d[None] for _, d in pending], axis=0)
        result = self.predict_fn(batch)
        self.inference_time += time.time() - t1
        lo = 0
        for c, d in pending:
            n = count(d)
            self._send(make_msg(_rows(result, lo, lo + n, d.shape == item_shape), cmd="predict_reply",
                                broker_id=c.get("broker_id", -1), explorer_id=c.get("explorer_id", -1)))
            lo += n
        self.iters += len(pending)
        self.batches += 1
        if self.iters > self._report_period:
            self._send(make_msg(self.stats(), cmd="stats_msg{}".format(self.predictor_name)))
        return len(pending)

    def process(self):
        """predictor.py:47-66: serve forever."""
        while True:
            self.process_once(None)


class RemotePredict(object):
    """Explorer-side stub (cartpole_dqn.py:51-55): ``predict(state)`` = send the observation(s), block on the reply."""

    def __init__(self, send_q, recv_q, explorer_id=-1, broker_id=-1):
        self.send_q, self.recv_q = send_q, recv_q
        self.explorer_id, self.broker_id = explorer_id, broker_id

    def predict(self, state):
        msg = make_msg(state, cmd="predict", sub_cmd="predict", explorer_id=self.explorer_id, broker_id=self.broker_id)
        (self.send_q.send if hasattr(self.send_q, "send") else self.send_q.put)(msg)
        reply = self.recv_q.recv() if hasattr(self.recv_q, "recv") else self.recv_q.get()
        if isinstance(reply, dict) and "ctr_info" in reply:
            return reply["data"]
        if isinstance(reply, tuple) and len(reply) == 2 and isinstance(reply[0], dict):     # UniComm.recv(): (ctr_info, data)
            return reply[1]
        return reply

    def sync_weights(self, weights):
        msg = make_msg(weights, cmd="predict", sub_cmd="sync_weights", explorer_id=self.explorer_id, broker_id=self.broker_id)
        (self.send_q.send if hasattr(self.send_q, "send") else self.send_q.put)(msg)
