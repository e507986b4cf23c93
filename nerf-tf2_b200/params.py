"""
Configuration object with the attribute layout of the reference's python-box `params`
(utils/params_utils.py:4-13; keys from params/config.yaml). Only the keys the hot path reads are
given defaults: sampling.{N_coarse,N_fine,perturb,lin_inv_depth} (config.yaml:364-381),
system.white_bg (:15), data.batch_size (:161). `load_params` accepts a YAML path or a dict.
"""
import copy
import types

DEFAULTS = {
    "system": {"white_bg": False, "run_eagerly": False, "log_images": False, "tf_seed": 11},
    "data": {"batch_size": 4096},
    "sampling": {"N_coarse": 64, "N_fine": 128, "perturb": True, "lin_inv_depth": True},
    "model": {"load": {"set_weights": False}},
}


def _to_ns(d):
    if isinstance(d, dict):
        return types.SimpleNamespace(**{k: _to_ns(v) for k, v in d.items()})
    return d


def _merge(base, over):
    out = copy.deepcopy(base)
    for k, v in (over or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def make_params(overrides=None, **sampling):
    """make_params({"system": {"white_bg": True}}, N_fine=256) -> attribute-access params."""
    d = _merge(DEFAULTS, overrides)
    d["sampling"].update(sampling)
    return _to_ns(d)


def load_params(path_or_dict):
    """utils/params_utils.load_params: YAML file (or a dict) -> attribute-access object."""
    if isinstance(path_or_dict, dict):
        return make_params(path_or_dict)
    import yaml
    with open(path_or_dict, "r") as f:
        return make_params(yaml.safe_load(f))
