import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O


def load_replays(name):
    import numpy as np

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    step_off = np.concatenate([[0], np.cumsum(d["n_steps"])])
    legal_off = np.concatenate([[0], np.cumsum(d["legal_len"])])
    games = []
    for g in range(len(d["P"])):
        lo, hi = step_off[g], step_off[g + 1]
        games.append(dict(
            P=int(d["P"][g]), limit=bool(d["limit"][g]), policy=str(d["policy"][g]), decks=d["decks"][g],
            nobles=d["nobles"][g][:int(d["P"][g]) + 1], seat=d["seat"][lo:hi], action=d["action"][lo:hi],
            reward=d["reward"][lo:hi], ends_plain=d["ends_plain"][lo:hi], ends_limit=d["ends_limit"][lo:hi],
            legal=[d["legal_flat"][legal_off[i]:legal_off[i + 1]] for i in range(lo, hi)],
            snap0=d["snap0"][g], snap=d["snap"][lo:hi], obs=d["obs"][lo:hi], score2=d["score2"][g][:int(d["P"][g])],
            final_obs=d["final_obs"][g][:int(d["P"][g])],
        ))
    return games


def load_env_episodes(name):
    """tests/golden/env_*.npz (oracle/gen_golden_env.py: the reference's own SplendorEnv driven call by call)."""
    import numpy as np

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    call_off = np.concatenate([[0], np.cumsum(d["n_calls"])])
    legal_off = np.concatenate([[0], np.cumsum(d["legal_len"])])
    all_off = np.concatenate([[0], np.cumsum(d["n_all"])])
    eps = []
    for g in range(len(d["P"])):
        lo, hi, P = call_off[g], call_off[g + 1], int(d["P"][g])
        eps.append(dict(
            P=P, my_id=int(d["my_id"][g]), seed=int(d["seed"][g]), game=int(d["game"][g]), policy=str(d["policy"][g]),
            opponents=str(d["opponents"][g]), decks=d["decks"][g], nobles=d["nobles"][g][:P + 1], action=d["action"][lo:hi],
            reward=d["reward"][lo:hi], terminated=d["terminated"][lo:hi], obs=d["obs"][lo:hi], snap=d["snap"][lo:hi],
            legal=[d["legal_flat"][legal_off[i]:legal_off[i + 1]] for i in range(lo, hi)],
            all_actions=d["all_actions"][all_off[g]:all_off[g + 1]], pad=int(d["pad"][g]) if "pad" in d else 0,
        ))
    return eps


def scripted_choice(idxs, t):
    """The deterministic hoarding policy of oracle/gen_golden_env.py's scripted episodes (same definition)."""
    pool = [i for i in idxs if 510 <= i < 3420] or [i for i in idxs if 6 <= i < 510] or list(idxs)
    return pool[(7 * t + 3) % len(pool)]
