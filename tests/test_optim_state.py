"""Checkpoint contract of the optimizer (utils/net_utils.py:5-40 saves / restores ``optim.state_dict()``): the
state_dict of the B200 Adam must be loadable by a stock ``torch.optim.Adam`` over the same parameters and vice
versa.  Host logic only (the update kernel itself is checked on the GPU in tests/test_gpu_resume.py)."""
import pytest
import torch


def _host_adam(params, step=3, lr=1e-3):
    """An optim.Adam whose tensors live on the CPU: __init__ refuses CPU parameters (no CPU compute path), the
    state_dict plumbing under test never launches a kernel."""
    from selectivenet_for_semantic_segmentation_binary_b200.optim import Adam
    opt = Adam.__new__(Adam)
    opt.params = list(params)
    opt.device = torch.device("cpu")
    opt.defaults = dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
    opt.param_groups = [dict(opt.defaults, params=opt.params)]
    g = torch.Generator().manual_seed(1)
    opt.exp_avg = [torch.randn(p.shape, generator=g) for p in opt.params]
    opt.exp_avg_sq = [torch.rand(p.shape, generator=g) for p in opt.params]
    opt.lr_dev = torch.tensor([lr])
    opt.step_dev = torch.tensor([step], dtype=torch.int32)
    opt._lr_cached = lr
    return opt


def test_state_dict_has_torch_adam_layout_and_loads_into_torch():
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ours = _host_adam(params)
    sd = ours.state_dict()
    ref = torch.optim.Adam(params, lr=1e-3)
    for p in params:
        p.grad = torch.ones_like(p)
    ref.step()
    ref_sd = ref.state_dict()
    assert set(sd) == set(ref_sd) == {"state", "param_groups"}
    assert set(sd["state"][0]) == set(ref_sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert set(ref_sd["param_groups"][0]) <= set(sd["param_groups"][0])
    assert sd["param_groups"][0]["params"] == [0, 1]
    ref.load_state_dict(sd)                               # what utils/net_utils.net_train_load does
    st = ref.state_dict()["state"]
    assert float(st[0]["step"]) == 3.0 and torch.equal(st[1]["exp_avg"], ours.exp_avg[1])
    assert torch.equal(st[0]["exp_avg_sq"], ours.exp_avg_sq[0])


def test_load_state_dict_accepts_torch_adam_and_rejects_stubs():
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ref = torch.optim.Adam(params, lr=2e-3, weight_decay=5e-4)
    for _ in range(2):
        for p in params:
            p.grad = torch.randn_like(p)
        ref.step()
    ours = _host_adam(params, step=0)
    ours.load_state_dict(ref.state_dict())
    assert int(ours.step_dev) == 2 and ours.param_groups[0]["lr"] == 2e-3 and ours.param_groups[0]["weight_decay"] == 5e-4
    assert float(ours.lr_dev) == pytest.approx(2e-3)
    rs = ref.state_dict()["state"]
    for i in range(2):
        assert torch.equal(ours.exp_avg[i], rs[i]["exp_avg"]) and torch.equal(ours.exp_avg_sq[i], rs[i]["exp_avg_sq"])
    with pytest.raises(ValueError):                       # the round-1 stub {'step','lr','type'} is not an Adam state
        ours.load_state_dict({"step": 3, "lr": 1e-3, "type": "sunet_b200.Adam"})
    with pytest.raises(ValueError):
        ours.load_state_dict({"state": {}, "param_groups": [dict(ref.state_dict()["param_groups"][0], params=[0])]})
    # a fresh optimizer (no step taken yet) round-trips as an empty state
    fresh = _host_adam(params, step=0)
    assert fresh.state_dict()["state"] == {}
    ours.load_state_dict(fresh.state_dict())
    assert int(ours.step_dev) == 0 and float(ours.exp_avg[0].abs().sum()) == 0.0


def test_shard_bounds_never_leaves_a_rank_empty():
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import chunk_bounds, shard_bounds
    for total, world in ((128, 8), (128, 2), (16, 3), (100, 8), (9, 8), (17, 8), (8, 8), (5, 2)):
        b = [shard_bounds(total, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == total
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1)) and all(hi > lo for lo, hi in b)
        chunk = [chunk_bounds(total, world, r) for r in range(world)]
        if all(hi > lo for lo, hi in chunk):               # torch.chunk sizes whenever they are usable
            assert b == chunk
    with pytest.raises(ValueError):
        shard_bounds(3, 8, 0)
