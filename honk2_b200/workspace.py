"""Checkpoint compatibility with the reference's ``Workspace`` (/root/reference/utils/workspace.py:28-70).

The reference saves ``{"model_state_dict", "loss_fn", "metrics", "optimizer_state_dict", "lr_scheduler_state_dict",
"epoch", ...}`` with ``torch.save`` (:28-45); ``loss_fn`` and ``metrics`` are pickled Python objects, so the file
needs ``weights_only=False``, and a model saved from under ``torch.nn.DataParallel`` (run/test.py:69-70) has a
``module.`` prefix on every key.  ``load_checkpoint`` puts the weights into a honk2_b200 model (strict, like
``Workspace._load`` :58-61; the packed device copies are rebuilt on the next forward) and returns the rest.
"""
import torch


def strip_data_parallel_prefix(state_dict):
    """``module.layers.conv_0.weight`` -> ``layers.conv_0.weight`` (only when EVERY key has the prefix)."""
    keys = list(state_dict.keys())
    if keys and all(k.startswith("module.") for k in keys):
        return type(state_dict)((k[len("module."):], v) for k, v in state_dict.items())
    return state_dict


def load_checkpoint(model, path_or_dict, map_location="cpu"):
    """Load a reference ``checkpoint_N.pt`` / ``best_model.pt`` (or an already loaded dict, or a bare state_dict)
    into ``model``; returns the remaining entries (epoch, metrics, ...) like ``Workspace._load`` does."""
    ckpt = path_or_dict
    if not isinstance(ckpt, dict):
        ckpt = torch.load(path_or_dict, map_location=map_location, weights_only=False)
    ckpt = dict(ckpt)
    sd = ckpt.pop("model_state_dict", None)
    if sd is None:          # a bare state_dict
        sd, ckpt = ckpt, {}
    model.load_state_dict(strip_data_parallel_prefix(sd))   # strict, workspace.py:61
    ckpt.pop("optimizer_state_dict", None)
    ckpt.pop("lr_scheduler_state_dict", None)
    return ckpt
