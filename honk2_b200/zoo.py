"""Model zoo: the ``model`` blocks of honk2's shipped configs (hyper-parameters only), keyed by
config name, so that benches/tests do not need /root/reference at run time.

Source of each entry: /root/reference/config/resnet/*.json, config/cnn/*.json,
config/hey_snips/res26.json, config/gsc_dev_config.json (``"model"`` block, lines 3-15 or
3-50).  ``n_labels`` is what run/test.py:54-58 injects (targets + unknown + silence); ``seed``
is the config's ``seed``.
"""
import copy


def _cnn(conv_0, pool_0, conv_1=None, pool_1=None, tail=("lin_0", "dnn_0", "dnn_1")):
    cfg = {"time": 101, "frequency": 40, "dropout_prob": 0.5,
           "conv_0": {"out_channels": conv_0[0], "kernel_size": list(conv_0[1]), "stride": list(conv_0[2])},
           "pool_0": {"kernel_size": list(pool_0)}}
    if conv_1 is not None:
        cfg["conv_1"] = {"out_channels": conv_1[0], "kernel_size": list(conv_1[1]), "stride": [1, 1]}
        cfg["pool_1"] = {"kernel_size": list(pool_1)}
    widths = {"lin_0": 32, "dnn_0": 128, "dnn_1": 128}
    for name in tail:
        cfg[name] = {"out_features": widths[name]}
    return cfg


def _res(n_layers, n_maps, use_dilation, pool=None, **extra):
    cfg = {"n_feature_maps": n_maps, "n_layers": n_layers, "use_dilation": use_dilation}
    if pool is not None:
        cfg = {"pool": list(pool), **cfg}
    cfg.update(extra)
    return cfg


MODEL_ZOO = {
    # config/resnet/*.json
    "res8":         {"name": "ResNet", "n_labels": 12, "seed": 0,   "config": _res(6, 45, False, (4, 3))},
    "res8_narrow":  {"name": "ResNet", "n_labels": 12, "seed": 100, "config": _res(6, 19, False, (4, 3))},
    "res15":        {"name": "ResNet", "n_labels": 12, "seed": 0,   "config": _res(13, 45, True)},
    "res15_narrow": {"name": "ResNet", "n_labels": 12, "seed": 100, "config": _res(13, 19, True)},
    "res26":        {"name": "ResNet", "n_labels": 12, "seed": 0,   "config": _res(24, 45, False, (2, 2))},
    "res26_narrow": {"name": "ResNet", "n_labels": 12, "seed": 100, "config": _res(24, 19, False, (2, 2))},
    # config/hey_snips/res26.json:5-13 -- the pooling key is spelt "avg_pool", which
    # resnet.py:29 ("pool" in config) never reads: 24 dilated layers (<=128) on the full map.
    "hey_snips_res26": {"name": "ResNet", "n_labels": 2, "seed": 0,
                        "config": _res(24, 45, True, None, avg_pool=[2, 2])},
    # config/gsc_dev_config.json:4-15 == config/hey_snips_dev_config.json:4-15
    "dev":          {"name": "ResNet", "n_labels": 12, "seed": 100, "config": _res(6, 19, False, (4, 3))},
    # config/cnn/*.json
    "cnn-trad-fpool3":  {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((64, (20, 8), (1, 1)), (1, 3), (64, (10, 4)), (1, 1), ("lin_0", "dnn_0"))},
    "cnn-trad-pool2":   {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((64, (20, 8), (1, 1)), (2, 2), (64, (10, 4)), (1, 1), ())},
    "cnn-one-fpool3":   {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((54, (32, 8), (1, 1)), (1, 3))},
    "cnn-one-fstride4": {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((186, (32, 8), (1, 4)), (1, 3))},
    "cnn-one-fstride8": {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((336, (32, 8), (1, 8)), (1, 3))},
    "cnn-tstride2":     {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((78, (16, 8), (2, 1)), (1, 3), (78, (9, 4)), (1, 1))},
    "cnn-tstride4":     {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((100, (16, 8), (4, 1)), (1, 3), (78, (5, 4)), (1, 1))},
    "cnn-tstride8":     {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((126, (16, 8), (8, 1)), (1, 3), (78, (5, 4)), (1, 1))},
    "cnn-tpool2":       {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((94, (21, 8), (1, 1)), (2, 3), (94, (6, 4)), (1, 1))},
    "cnn-tpool3":       {"name": "CNN", "n_labels": 12, "seed": 0,
                         "config": _cnn((94, (15, 8), (1, 1)), (3, 3), (94, (6, 4)), (1, 1))},
}


def model_config(name, n_labels=None):
    """Return (class_name, config_dict) exactly as run/test.py:60-64 would pass them: the
    ``model.config`` block with ``n_labels`` injected."""
    entry = MODEL_ZOO[name]
    cfg = copy.deepcopy(entry["config"])
    cfg["n_labels"] = entry["n_labels"] if n_labels is None else n_labels
    return entry["name"], cfg
