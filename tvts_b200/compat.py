"""Small host-side helpers the reference model constructors use (restated; v2/utils/util.py:25-51)."""
from collections import OrderedDict


def state_dict_data_parallel_fix(load_state_dict, curr_state_dict):
    """Add / strip the DataParallel `module.` prefix so `load_state_dict` matches `curr_state_dict`'s naming."""
    load_keys, curr_keys = list(load_state_dict.keys()), list(curr_state_dict.keys())
    cur_dp = curr_keys[0].startswith("module.")
    load_dp = load_keys[0].startswith("module.")
    if load_dp and not cur_dp:
        return OrderedDict((k[7:], v) for k, v in load_state_dict.items())
    if cur_dp and not load_dp:
        return OrderedDict(("module." + k, v) for k, v in load_state_dict.items())
    return load_state_dict
