"""Small host helpers with the reference's names (ssr_eval/utils.py): aggregation and JSON I/O.
The numeric helpers of the reference's utils (to_log / pow_p_norm / energy_unify / pow_norm) have no
host-side counterpart here: their arithmetic lives inside the fused CUDA kernel."""
import json
import wave

import numpy as np

EPS = 1e-12  # ssr_eval/utils.py:7


def dict_mean(dict_list):
    """Per-key float64 mean of a list of dicts (ssr_eval/utils.py:24-28)."""
    first = dict_list[0]
    return {k: np.mean([d[k] for d in dict_list]) for k in first.keys()}


def write_json(obj, fname):
    """ssr_eval/utils.py:18-21 (indent 4)."""
    with open(fname, "w") as f:
        f.write(json.dumps(obj, indent=4))


def load_json(fname):
    with open(fname, "r") as f:
        return json.load(f)


def get_sample_rate(fname):
    with wave.open(fname) as f:
        return f.getparams()[2]


def get_framesLength(fname):
    with wave.open(fname) as f:
        return f.getparams()[3]


def write_list(items, fname):
    with open(fname, "w") as f:
        for w in items:
            f.write(w + "\n")


def read_list(fname):
    with open(fname, "r") as f:
        return [line.strip("\n") for line in f.readlines()]
