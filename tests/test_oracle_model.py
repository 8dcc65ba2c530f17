"""Pins oracle/model_ref.py (the functional restatement of ResNet.forward / CNN.forward) against
logits produced by the UNMODIFIED reference modules (committed in tests/golden/model_golden.npz
by oracle/make_golden.py) and, when /root/reference is present, against the live modules."""
import numpy as np
import pytest
import torch

import honk2_b200
from honk2_b200 import synth
from honk2_b200.zoo import MODEL_ZOO, model_config
from oracle import model_ref, reference_loader
from oracle.make_golden import weight_checksum

ZOO = list(MODEL_ZOO)


def build_state(name, variant):
    """Our module under the config's seed -> the same default-init parameters the reference
    module gets (same constructor order); optionally hardened."""
    m = honk2_b200.build_model(name)
    sd = m.state_dict()
    if variant == "hardened":
        synth.harden_(sd)
    return m, sd


@pytest.mark.parametrize("name", ZOO)
@pytest.mark.parametrize("variant", ["default", "hardened"])
def test_restatement_matches_reference_golden(name, variant, model_golden):
    kind, cfg = model_config(name)
    _, sd = build_state(name, variant)
    assert weight_checksum(sd) == pytest.approx(float(model_golden[f"{name}/{variant}/wsum"]), rel=1e-12), \
        "our constructor does not reproduce the reference's default initialisation"
    x = torch.from_numpy(model_golden["feats"])
    y = model_ref.forward(kind, sd, cfg, x).numpy()
    ref = model_golden[f"{name}/{variant}/logits"]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    assert np.array_equal(y.argmax(1), ref.argmax(1))


@pytest.mark.parametrize("name", [n for n in ZOO if MODEL_ZOO[n]["name"] == "ResNet"])
def test_resnet_free_time_axis(name, model_golden):
    """ResNet accepts any T (global mean, resnet.py:57-58)."""
    kind, cfg = model_config(name)
    _, sd = build_state(name, "hardened")
    y = model_ref.forward(kind, sd, cfg, torch.from_numpy(model_golden["feats_long"])).numpy()
    ref = model_golden[f"{name}/hardened/logits_long"]
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.skipif(not reference_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("name", ["res8", "res15_narrow", "cnn-trad-fpool3", "cnn-tstride4"])
def test_restatement_matches_live_reference(name):
    kind, cfg = model_config(name)
    ref = reference_loader.build_model(kind, cfg, MODEL_ZOO[name]["seed"])
    sd = ref.state_dict()
    synth.harden_(sd, seed=77)
    x = torch.randn(3, 101, 40, generator=torch.Generator().manual_seed(5)) * 4 - 10
    with torch.no_grad():
        want = ref(x)
    got = model_ref.forward(kind, sd, cfg, x)
    assert torch.allclose(got, want, rtol=0, atol=1e-5)
    # state_dict keys of our module are the reference's (utils/workspace.py:61 loads strictly)
    ours = honk2_b200.build_model(name)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict(), strict=True)


def test_acc_counts():
    logits = torch.tensor([[0.1, 0.9], [0.8, 0.2], [0.5, 0.5]])
    assert model_ref.acc_counts(logits, torch.tensor([1, 1, 0])) == (2, 3)   # tie -> index 0
