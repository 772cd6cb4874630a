"""Host logic of MagCache (kandinsky/magcache_utils.py mirror) on CPU: the skip schedule for the reference's own
calibration curves equals what the reference's state machine produced when the golden was minted, and the schedule
resets after a full sample."""
import os

import numpy as np
import torch

from kandinsky.magcache_utils import MagCacheState, nearest_interp

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_skip_schedule_matches_reference_golden():
    rec = torch.load(os.path.join(GOLD, "tiny_sampler_magcache.pt"), weights_only=False)
    st = MagCacheState(rec["mag_ratios"], rec["steps"], no_cfg=False)
    got = [st.next() for _ in range(2 * rec["steps"])]
    assert [skip for _, skip in got] == rec["skipped"]
    assert st.cnt == 0 and st.accumulated_err == [0.0, 0.0]
    again = [st.next() for _ in range(2 * rec["steps"])]
    assert again == got                                               # second sample: same schedule


def test_no_cfg_walks_the_conditional_curve_only():
    ratios = list(np.linspace(0.95, 1.05, 18))
    st = MagCacheState(ratios, 10, no_cfg=True)
    slots = [st.next()[0] for _ in range(10)]
    assert slots == [0] * 10 and st.cnt == 0


def test_first_fifth_of_the_schedule_never_skips_and_runs_are_bounded_by_K():
    ratios = [1.0] * 98                                               # perfectly flat curve: error stays 0
    st = MagCacheState(ratios, 50, no_cfg=False)
    skips = [st.next()[1] for _ in range(100)]
    assert not any(skips[:20])                                        # retention_ratio 0.2 of 100 forwards
    cond = skips[20::2]
    run = best = 0
    for s in cond:
        run = run + 1 if s else 0
        best = max(best, run)
    assert best == MagCacheState.K                                    # at most K consecutive skips per branch


def test_nearest_interp_resamples_a_curve():
    src = np.arange(10.0)
    assert list(nearest_interp(src, 10)) == list(src)
    assert list(nearest_interp(src, 1)) == [9.0]
    out = nearest_interp(src, 4)
    assert out[0] == 0.0 and out[-1] == 9.0 and len(out) == 4
    st = MagCacheState(list(np.linspace(0.9, 1.1, 98)), 25, no_cfg=False)   # 100-entry curve resampled to 50 forwards
    assert len(st.mag_ratios) == 50
