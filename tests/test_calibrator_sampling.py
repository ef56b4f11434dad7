"""CPU test of the calibrator's opt-in "reference" mode (backend.Calibrator.replacement_indices): the seeded
emulation of calibrator.cc:6-23 — a 1000-slot sample where every later value overwrites a uniformly chosen slot
with probability 1000 / 2001 — against a literal restatement of that loop. The reference seeds its mt19937 from
random_device, so only DISTRIBUTIONS can be compared: which stream positions the slots hold at the end."""
import numpy as np

from int8inferenceengine_b200.backend import NUM_SAMPLES, Calibrator


def literal_last_writer(n, rng):
    """calibrator.cc:13-21 verbatim for one call of n values into an empty calibrator: index of the value each slot
    holds at the end."""
    owner = np.full(NUM_SAMPLES, -1, np.int64)
    cnt = 0
    draws = rng.integers(0, 2 * NUM_SAMPLES + 1, size=n)     # uniform_int_distribution(0, num_samples * 2), inclusive
    for i in range(n):
        if cnt < NUM_SAMPLES:
            owner[cnt] = i
            cnt += 1
        elif draws[i] < NUM_SAMPLES:
            owner[draws[i]] = i
    return owner


def emulated_last_writer(n, rng):
    owner = np.arange(NUM_SAMPLES, dtype=np.int64)            # the first 1000 values fill the slots in order
    idx = Calibrator.replacement_indices(n, NUM_SAMPLES, rng)
    return np.where(idx >= 0, idx, owner)


def test_replacement_indices_match_the_reference_process_in_distribution():
    n = 40000
    lit = np.concatenate([literal_last_writer(n, np.random.default_rng(s)) for s in range(6)])
    emu = np.concatenate([emulated_last_writer(n, np.random.default_rng(100 + s)) for s in range(6)])
    back_lit, back_emu = (n - 1) - lit, (n - 1) - emu          # distance from the end of the stream
    # geometric with mean 2000: compare mean, median and the far tail of 6000 slots each
    assert abs(back_lit.mean() - 2000) < 150 and abs(back_emu.mean() - 2000) < 150
    for q in (0.25, 0.5, 0.9, 0.99):
        a, b = np.quantile(back_lit, q), np.quantile(back_emu, q)
        assert abs(a - b) <= 0.12 * max(a, b) + 20, (q, a, b)
    # every slot holds a value of the stream; slots that were never hit again keep their initial value
    assert emu.min() >= 0 and emu.max() <= n - 1
    never_lit, never_emu = (lit < NUM_SAMPLES).mean(), (emu < NUM_SAMPLES).mean()
    assert never_lit < 1e-3 and never_emu < 1e-3               # (1 - 1/2001)^39000 ~ 3e-9


def test_replacement_indices_short_streams_keep_the_head():
    rng = np.random.default_rng(0)
    idx = Calibrator.replacement_indices(NUM_SAMPLES, NUM_SAMPLES, rng)      # nothing after the head
    assert (idx == -1).all()
    idx = Calibrator.replacement_indices(NUM_SAMPLES + 50, NUM_SAMPLES, rng)
    assert ((idx == -1) | ((idx >= NUM_SAMPLES) & (idx < NUM_SAMPLES + 50))).all()
    assert (idx >= 0).sum() < 60     # 50 values, each hits some slot with probability 1000/2001
