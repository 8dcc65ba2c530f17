"""CPU ORACLE (test infrastructure, NOT product code) -- plain PyTorch fp32 restatement of
the forward pass of honk2's registered models, driven by a reference-keyed ``state_dict``.

  resnet_forward  follows /root/reference/model/resnet.py:38-60  (construction :11-36)
  cnn_forward     follows /root/reference/model/cnn.py:79-107    (construction :12-77)

PARITY PIN: checked bit-for-bit against the reference modules themselves (imported from
/root/reference by oracle/reference_loader.py, in the build container only) in
tests/test_oracle_model.py, and against tests/golden/model_*.npz that oracle/make_golden.py
generated from those reference modules.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, resnet.py:27


def resnet_dilation(i, use_dilation):
    """resnet.py:21-26: dilation == padding == 2**((i-1)//3) when use_dilation else 1."""
    return int(2 ** ((i - 1) // 3)) if use_dilation else 1


def resnet_forward(sd, config, x, return_pooled=False):
    """x: [B, T, F] float32 -> logits [B, n_labels] (resnet.py:38-60).  `return_pooled` also
    returns the [B, C] input of the output layer (used to calibrate diverse-argmax test weights)."""
    n_layers = config["n_layers"]
    x = x.unsqueeze(1)                                                  # :39
    x = F.relu(F.conv2d(x, sd["layers.conv_0.weight"], padding=1))      # :40-41
    if "pool" in config:                                                # :43-44 (and :29)
        x = F.avg_pool2d(x, tuple(config["pool"]))
    prev_x = x                                                          # :46
    for i in range(1, n_layers + 1):
        d = resnet_dilation(i, config["use_dilation"])
        x = F.relu(F.conv2d(x, sd[f"layers.conv_{i}.weight"], padding=d, dilation=d))  # :48-49
        if i % 2 == 0:                                                  # :51-53
            x = x + prev_x
            prev_x = x
        x = F.batch_norm(x, sd[f"layers.bn_{i}.running_mean"], sd[f"layers.bn_{i}.running_var"],
                         None, None, False, 0.0, BN_EPS)                # :55 (eval, affine=False)
    x = x.view(x.size(0), x.size(1), -1).mean(2)                        # :57-58
    y = F.linear(x, sd["layers.output.weight"], sd["layers.output.bias"])  # :59
    return (y, x) if return_pooled else y


def cnn_forward(sd, config, x):
    """x: [B, T, F] float32 -> logits [B, n_labels] (cnn.py:79-107); dropout is identity in
    eval mode (run/test.py:21)."""
    x = x.unsqueeze(1)                                                  # :80
    x = F.conv2d(x, sd["layers.conv_0.weight"], sd["layers.conv_0.bias"],
                 stride=tuple(config["conv_0"]["stride"]))              # :82
    x = F.relu(x)                                                       # :83
    x = F.max_pool2d(x, tuple(config["pool_0"]["kernel_size"]))         # :85
    if "conv_1" in config:                                              # :87-91
        x = F.conv2d(x, sd["layers.conv_1.weight"], sd["layers.conv_1.bias"],
                     stride=tuple(config["conv_1"]["stride"]))
        x = F.relu(x)
        x = F.max_pool2d(x, tuple(config["pool_1"]["kernel_size"]))
    x = x.reshape(x.size(0), -1)                                        # :93
    for name in ("lin_0", "dnn_0", "dnn_1"):                            # :95-104
        if name in config:
            x = F.linear(x, sd[f"layers.{name}.weight"], sd[f"layers.{name}.bias"])
    return F.linear(x, sd["layers.lin_1.weight"], sd["layers.lin_1.bias"])  # :106


def forward(kind, sd, config, x):
    with torch.no_grad():
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items() if v.is_floating_point()}
        x = x.detach().to("cpu", torch.float32)
        if kind == "ResNet":
            return resnet_forward(sd, config, x)
        if kind == "CNN":
            return cnn_forward(sd, config, x)
        raise KeyError(kind)


def acc_counts(logits, target):
    """metric/acc.py:14-22: (correct, total) from argmax(dim=1) == target."""
    pred = torch.argmax(logits, dim=1)
    return int(torch.sum(pred == target).item()), int(len(target))


def per_class_counts(logits, target):
    """metric/per_class_acc.py:14-45: {class: (total, correct)} over the classes present in `target`."""
    pred = torch.argmax(logits, dim=1).tolist()
    out = {}
    for guess, truth in zip(pred, target.tolist()):
        tot, cor = out.get(truth, (0, 0))
        out[truth] = (tot + 1, cor + (1 if guess == truth else 0))
    return out


def ce_loss(logits, target):
    """loss_function.py:7-9: nn.CrossEntropyLoss()(output, target) = mean over the batch of
    logsumexp(row) - row[target] (float64 here)."""
    x = logits.double()
    lse = torch.logsumexp(x, dim=1)
    return float((lse - x[torch.arange(x.shape[0]), target]).mean())
