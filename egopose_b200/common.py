"""core/common.py:5-25 estimate_advantages with the reference signature, on the GAE scan kernel."""
from . import lib


def estimate_advantages(rewards, masks, values, gamma, tau):
    """rewards [N], masks [N], values [N, 1] CUDA float64 tensors -> (advantages [N, 1] standardised with
    the unbiased std, returns [N, 1]); one egp_gae_f64 launch + one standardise launch."""
    adv, ret, stats = lib.gae(rewards.contiguous().view(-1), masks.contiguous().view(-1),
                              values.contiguous().view(-1), float(gamma), float(tau))
    lib.standardize_(adv, stats)
    return adv.view(-1, 1), ret.view(-1, 1)
