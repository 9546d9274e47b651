"""'auto' width helpers (ref: offpolicy_rnn/policy_value_models/utils.py:3-23)."""
import math


def nearest_power_of_two_half(x):
    e = max(round(math.log(0.5 * x, 2)), 0)
    return int(math.ceil(2 ** e))


def nearest_power_of_two(x):
    e = max(int(math.ceil(math.log(x, 2))), 0)
    return int(math.ceil(2 ** e))
