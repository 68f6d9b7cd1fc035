"""The reference's policy networks (model.py) as torch modules: shapes and layer inventory on the CPU."""
import pytest
import torch

from gym_pcgrl_b200.models import ActorCritic, Cnn1, Cnn2, FullyConv1, FullyConv2, policy_for


def test_policy_selection_follows_train_py():
    assert policy_for("binary", "narrow") == "CustomPolicyBigMap" and policy_for("sokoban", "turtle") == "CustomPolicySmallMap"
    assert policy_for("zelda", "wide") == "FullyConvPolicyBigMap" and policy_for("sokoban", "wide") == "FullyConvPolicySmallMap"


def test_cnn_extractors_shapes():
    x = torch.randint(0, 2, (5, 28, 28, 1), dtype=torch.uint8)
    assert Cnn2((28, 28, 1))(x).shape == (5, 512)          # 28 -> 13 -> 6 -> 4: fc over 4*4*64
    assert Cnn2((28, 28, 1)).fc1.in_features == 4 * 4 * 64
    assert Cnn1((10, 10, 5)).fc1.in_features == 4 * 4 * 64 and Cnn1((10, 10, 5))(torch.zeros(2, 10, 10, 5)).shape == (2, 512)
    with pytest.raises(ValueError):
        Cnn1((5, 5, 5))


def test_fully_conv_policies_shapes_and_logit_order():
    obs = torch.randint(0, 2, (3, 14, 14, 1), dtype=torch.uint8)
    net = ActorCritic("FullyConvPolicyBigMap", (14, 14, 1), 14 * 14 * 2)
    logits, value = net(obs)
    assert logits.shape == (3, 14 * 14 * 2) and value.shape == (3,)
    assert len(net.extractor.body) == 8 and len(net.extractor.value) == 3       # c1..c8, v1 v2 v4
    assert net.extractor.vf_features == 2 * 2 * 64                               # 14 -> 6 -> 2
    small = ActorCritic("FullyConvPolicySmallMap", (5, 5, 5), 5 * 5 * 5)
    lg, v = small(torch.zeros(2, 5, 5, 5))
    assert lg.shape == (2, 125) and len(small.extractor.value) == 2 and small.extractor.vf_features == 2 * 2 * 64
    # logits are flattened in (h, w, tool) order == ActionMap's unravel order: tool varies fastest
    act, _ = FullyConv1((5, 5, 5), 5)(torch.zeros(1, 5, 5, 5))
    assert act.shape == (1, 125)


def test_feedforward_policy_heads():
    net = ActorCritic("CustomPolicyBigMap", (28, 28, 1), 3)
    logits, value = net(torch.zeros(4, 28, 28, 1, dtype=torch.uint8))
    assert logits.shape == (4, 3) and value.shape == (4,)
    assert float(net.pi.weight.norm()) < 0.1          # init_scale 0.01 head
    assert all(float(b.abs().sum()) == 0 for n_, b in net.named_parameters() if n_.endswith("bias"))
