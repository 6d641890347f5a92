"""Helpers shared by the golden-fixture tests (CPU oracle and GPU parity)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EASY_FIXTURES = ["easy_1a_am0", "easy_3a_am0", "easy_3a_am1", "easy_5a_am2", "easy_5a_am3",
                 "easy_5a_crowded", "easy_3a_tm1"]
FLIGHT_FIXTURES = ["flight_3a_am0", "flight_2a_small"]
SEARCH_FIXTURES = ["search_3a_default", "search_4a_am1_tm1", "search_5a_am2", "search_64a_1000t"]


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def flight_spec_kwargs(g, variant):
    n, m, M, R, T, am, tm, base, seed = [int(v) for v in g["meta"]]
    vel, d, safe, fd = [float(v) for v in g["fmeta"]]
    as_num = lambda v: int(v) if float(v).is_integer() else v
    return dict(n_agents=n, target_num=m, map_size=M, view_range=R, time_limit=T, agent_mode=am,
                target_mode=tm, velocity=as_num(vel), detect_prob=d, safe_dist=as_num(safe),
                force_dist=as_num(fd), variant=variant), base, seed


TEMPLATE = {
    # flight_targets.txt of the reference (main.py:19-32 parsing), embedded so the GPU box needs no reference tree
    "x": [5, 2, 7.5, 2.8, 6.9, 5.5, 5.3, 1.8, 3, 4.5, 6.3, 8, 0.9, 9.4, 4.2],
    "y": [9.1, 7.5, 7, 8, 8.5, 8, 6.6, 6.8, 5.7, 5, 5.7, 6.7, 8.7, 9, 9.3],
    "deter": ["f", "t", "f", "f", "t", "f", "t", "t", "f", "f", "t", "f", "f", "t", "f"],
    "priority": [3, 3, 3, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1],
    "dx": [0.2, 0.3, 0.3, 0.27, 0.25, 0.25, 0.1, 0.28, 0.18, 0.23, 0.31, 0.29, 0.15, 0.21, 0.34],
    "dy": [0.2, 0.3, 0.26, 0.27, 0.25, 0.25, 0.12, 0.28, 0.18, 0.25, 0.30, 0.28, 0.16, 0.21, 0.33],
}


