"""Checkpoint bookkeeping helpers with the reference's file naming (metabo/ppo/util.py:43-74)."""
import os
import pickle as pkl

import numpy as np


def get_last_iter_idx(logpath):
    its = [int(f[len("stats_"):]) for f in os.listdir(logpath) if f.startswith("stats_")]
    return max(its) if its else -np.inf


def get_best_iter_idx(logpath):
    """iteration with the best average step reward; cross-checked against that iteration's own stats file"""
    last = get_last_iter_idx(logpath)
    with open(os.path.join(logpath, "stats_{:d}".format(last)), "rb") as f:
        stats = pkl.load(f)
    rews = np.asarray(stats["avg_step_rews"])
    best = int(np.argmax(rews))
    with open(os.path.join(logpath, "stats_{:d}".format(best)), "rb") as f:
        best_stats = pkl.load(f)
    assert best_stats["batch_stats"]["avg_step_reward"] == rews[best]
    return best
