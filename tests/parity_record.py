"""Observed parity errors of the GPU tests, written where the run artefacts travel back from the GPU box."""
import os


def record(name, payload):
    """Observed parity errors are written next to the other run artefacts (gpurun_out/ travels back from the GPU box)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "parity_errors.json")
        try:
            with open(path) as f:
                cur = json.load(f)
        except Exception:
            cur = {}
        cur[name] = payload
        with open(path, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
    print("parity[%s] %s" % (name, payload))
