"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of VecFrameStack's observation update.

Follows baselines/common/vec_env/vec_frame_stack.py:17-31.  Pinned against tests/golden/frame_stack_*.npz, which
oracle/gen_golden.py produced by executing the reference class itself on a scripted venv.
"""
import numpy as np


def frame_stack_reset(first_frames, nstack):
    """vec_frame_stack.py:27-31: zero stack, newest slot = first observation."""
    c = first_frames.shape[-1]
    stacked = np.zeros(first_frames.shape[:-1] + (nstack * c,), first_frames.dtype)
    stacked[..., -c:] = first_frames
    return stacked


def frame_stack_step(stacked, frames, news):
    """vec_frame_stack.py:17-25: roll the channel axis by -1 *element* (the reference shifts by one channel, which
    equals one frame only for single-channel frames -- restated literally), clear finished envs, insert frames."""
    out = np.roll(stacked, shift=-1, axis=-1)
    for i, new in enumerate(news):
        if new:
            out[i] = 0
    out[..., -frames.shape[-1]:] = frames
    return out
