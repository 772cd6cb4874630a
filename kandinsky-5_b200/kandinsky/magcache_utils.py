"""Drop-in for the reference's `kandinsky/magcache_utils.py` (an adaptation of MagCache): the same
`set_magcache_params(dit, mag_ratios, num_steps, no_cfg)` entry point and the same skip rule, but instead of swapping
the module's `forward` for a Python re-implementation of the block loop, the decision (pure host arithmetic on the
calibrated magnitude ratios) is kept here and the engine does the rest: `k5_dit_forward_magcache(slot, skip)` either
runs the 32 visual blocks and caches their residual, or replaces them by "embedded input + cached residual"."""
import numpy as np


def nearest_interp(src_array, target_length):
    """magcache_utils.py:6-13: nearest-neighbour resampling of a calibration curve to another step count."""
    src_array = np.asarray(src_array)
    if target_length == 1:
        return np.array([src_array[-1]])
    scale = (len(src_array) - 1) / (target_length - 1)
    return src_array[np.round(np.arange(target_length) * scale).astype(int)]


class MagCacheState:
    """The bookkeeping of magcache_utils.py:16-38 (set-up) and :64-80, :92-101 (per-forward decision).  `num_steps` counts
    sampler steps; the schedule is indexed by forward (2 per step: conditional, unconditional)."""

    thresh, K, retention_ratio = 0.12, 2, 0.2

    def __init__(self, mag_ratios, num_steps, no_cfg):
        self.num_forwards = num_steps * 2
        self.no_cfg = bool(no_cfg)
        ratios = np.array([1.0] * 2 + list(mag_ratios))
        if len(ratios) != self.num_forwards:
            con, ucon = nearest_interp(ratios[0::2], num_steps), nearest_interp(ratios[1::2], num_steps)
            ratios = np.concatenate([con.reshape(-1, 1), ucon.reshape(-1, 1)], axis=1).reshape(-1)
        self.mag_ratios = ratios
        self.reset()

    def reset(self):
        self.cnt = 0
        self.accumulated_err, self.accumulated_steps, self.accumulated_ratio = [0.0, 0.0], [0, 0], [1.0, 1.0]

    def next(self):
        """-> (slot, skip) for the forward about to run, and advance."""
        slot, skip = self.cnt % 2, False
        if self.cnt >= int(self.num_forwards * self.retention_ratio):
            self.accumulated_ratio[slot] = self.accumulated_ratio[slot] * self.mag_ratios[self.cnt]
            self.accumulated_steps[slot] += 1
            self.accumulated_err[slot] += np.abs(1 - self.accumulated_ratio[slot])
            if self.accumulated_err[slot] < self.thresh and self.accumulated_steps[slot] <= self.K:
                skip = True
            else:
                self.accumulated_err[slot], self.accumulated_steps[slot], self.accumulated_ratio[slot] = 0, 0, 1.0
        self.cnt += 2 if self.no_cfg else 1
        if self.cnt >= self.num_forwards:
            self.reset()
        return slot, skip


def set_magcache_params(dit, mag_ratios, num_steps, no_cfg):
    """Same call as the reference (kandinsky/utils.py:107-113).  Arms MagCache on the engine-backed DiT."""
    dit._magcache = MagCacheState(list(mag_ratios), int(num_steps), no_cfg)
    return dit
