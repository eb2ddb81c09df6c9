"""Training step on the GPU (include/alphagpu_train.h) against the CPU oracle: gradients, Adam moments and parameters bit for bit,
losses to 1e-6 relative.  Calls go through the C ABI (alphagpu_b200.train.Trainer is a ctypes shim)."""
import numpy as np
import pytest

import alphagpu_b200 as ag
import oracle
from helpers_train import make_batch, make_net

pytestmark = pytest.mark.gpu


def to_networkf(d):
    return ag.NetworkF(d["base"], d["res"], d["pol_w"], d["pol_b"], d["val_w"], d["val_b"], d["feat_w"], d["feat_b"])


def flat(orc, net):
    return orc.pack(dict(base=net.base, res=net.res, pol_w=net.policy, pol_b=net.policy_bias, val_w=net.value, val_b=net.value_bias,
                         feat_w=net.feature, feat_b=net.feature_bias))


SHAPES = [
    # in, n, k, A, FS, B
    (18, 16, 0, 9, 9, 7),            # no residual blocks, tiny
    (84, 128, 5, 7, 42, 300),        # Connect4 128x6, batch crossing a 256-sample slice with a ragged tail
    (84, 128, 5, 7, 42, 512),
    (162, 96, 2, 81, 81, 130),       # Gobang9 heads (NH = 163 > 2 column tiles), in not a multiple of 4
    (128, 512, 1, 65, 64, 257),      # Reversi8 width 512
]


@pytest.mark.parametrize("inp,n,k,A,FS,B", SHAPES)
def test_gradient_bit_exact(inp, n, k, A, FS, B):
    d = make_net(inp, n, k, A, FS, seed=B, bias_scale=0.05)
    batch = make_batch(inp, A, FS, B, seed=B + 1)
    orc = oracle.Trainer(inp, n, k, A, FS)
    orc.set_params(d)
    want_loss = orc.loss_grad(*batch)
    tr = ag.Trainer(inp, n, k, A, FS, max_batch=B)
    tr.set_params(to_networkf(d))
    got_loss = tr.loss_grad(*batch)
    got = flat(orc, tr.get_grads())
    want = orc.get(orc.GRADS)
    assert np.array_equal(got, want), f"max |diff| {np.abs(got - want).max()} at {np.argmax(np.abs(got - want))}"
    assert np.allclose(got_loss, want_loss, rtol=1e-6, atol=1e-7), (got_loss, want_loss)
    assert np.allclose(tr.lossTot(*batch), want_loss, rtol=1e-6, atol=1e-7)
    tr.close()


def test_three_steps_bit_exact_params_and_moments():
    inp, n, k, A, FS, B = 84, 128, 3, 7, 42, 384
    d = make_net(inp, n, k, A, FS, seed=11, bias_scale=0.05)
    orc = oracle.Trainer(inp, n, k, A, FS)
    orc.set_params(d)
    tr = ag.Trainer(inp, n, k, A, FS, max_batch=B)
    tr.set_params(to_networkf(d))
    for s in range(3):
        batch = make_batch(inp, A, FS, B, seed=20 + s)
        lo, lg = orc.step(*batch), tr.step(*batch)
        assert np.allclose(lg, lo, rtol=1e-6, atol=1e-7)
        assert np.array_equal(flat(orc, tr.get_params()), orc.get(orc.PARAMS)), s
        m, v, bp = tr.opt_state()
        assert np.array_equal(m, orc.get(orc.M)) and np.array_equal(v, orc.get(orc.V)), s
        assert np.allclose(bp, [0.9 ** (s + 2), 0.999 ** (s + 2)], rtol=1e-14)
    # smaller batch than max_batch afterwards, and reset_optimizer
    batch = make_batch(inp, A, FS, 100, seed=99)
    orc.step(*batch); tr.step(*batch)
    assert np.array_equal(flat(orc, tr.get_params()), orc.get(orc.PARAMS))
    tr.close()


def test_params_roundtrip_and_errors():
    inp, n, k, A, FS = 18, 32, 2, 9, 9
    net = ag.ressimplesf_full(inp, A, FS, n, k, seed=4)
    net.policy_bias[:] = np.arange(A); net.feature_bias[:] = -np.arange(FS); net.value_bias[:] = 7
    tr = ag.Trainer.for_network(net, 64)
    back = tr.get_params()
    assert all(np.array_equal(a, b) for a, b in zip(back.arrays(), net.arrays()))
    with pytest.raises(ag._lib.AlphaGPUError) as e:
        tr.apply()                                         # no gradient yet
    assert e.value.code == ag._lib.ERR_STATE
    batch = make_batch(inp, A, FS, 65, seed=1)
    with pytest.raises(ag._lib.AlphaGPUError) as e:
        tr.step(*batch)                                    # B > max_batch
    assert e.value.code == ag._lib.ERR_INVALID
    with pytest.raises(ValueError):
        tr.step(batch[0][:, :-1], *batch[1:])
    m, v, bp = tr.opt_state()
    m[:] = 1.5; v[:] = 2.5
    tr.set_opt_state(m, v, [0.5, 0.25])
    m2, v2, bp2 = tr.opt_state()
    assert np.all(m2 == 1.5) and np.all(v2 == 2.5) and list(bp2) == [0.5, 0.25]
    tr.close()


def test_grad_tensor_aliases_device_gradient():
    import torch
    inp, n, k, A, FS, B = 18, 32, 1, 9, 9, 40
    net = ag.ressimplesf_full(inp, A, FS, n, k, seed=5)
    tr = ag.Trainer.for_network(net, B)
    batch = make_batch(inp, A, FS, B, seed=2)
    tr.loss_grad(*batch)
    g = tr.grad_tensor()
    orc = oracle.Trainer(inp, n, k, A, FS)
    assert g.is_cuda and g.numel() == orc.P
    assert np.array_equal(g.cpu().numpy(), flat(orc, tr.get_grads()))
    # what step_dp does at world 2 with identical shards: sum then scale 1/2 == the plain step
    before = flat(orc, tr.get_params())
    g.mul_(2.0)
    torch.cuda.synchronize()
    tr.apply(0.5)
    tr2 = ag.Trainer.for_network(net, B)
    tr2.step(*batch)
    assert np.array_equal(flat(orc, tr.get_params()), flat(orc, tr2.get_params()))
    assert not np.array_equal(before, flat(orc, tr.get_params()))
    tr.close(); tr2.close()


def test_generation_selfplay_train_duel_ttt(tmp_path):
    """trainingPipeline (selfplay.jl:1-109) end to end on tic-tac-toe: samples from the GPU self-play feed traininPipe, the trained
    net is handed back to the search through convert_back, the duel produces a finite Elo, the checkpoint reloads."""
    spec = ag.GameSpec.named("gobang", 3, 3)
    net = ag.ressimplesf_full(2 * spec.VectorizedState, spec.maxActions, spec.FeatureSize, 128, 2, seed=0)
    trainingnet = net.copy()
    buf = ag.PoolSample(spec, 50000)
    net2, tn2, passing, elo = ag.trainingPipeline(net, trainingnet, buf, 1, -1000.0, spec=spec, game="ttt", cpuct=1.5, samplesNumber=512, rollout=16,
                                                  batchsize=256, duel_games=64, duel_rollout=8, save_dir=str(tmp_path), verbose=False)
    assert buf.length_buffer() >= 512 * 5
    assert not np.array_equal(tn2.base, net.base)                      # trained
    assert np.isfinite(elo) or elo in (np.inf, -np.inf)
    assert passing == (elo > -1000.0)
    loaded, meta = ag.load_network(str(tmp_path / "reseau1.agpu"))
    assert meta["generation"] == 1 and all(np.array_equal(a, b) for a, b in zip(loaded.arrays(), tn2.arrays()))


def test_traininpipe_lowers_the_loss_on_selfplay_samples():
    spec = ag.GameSpec.named("connect4")
    net = ag.ressimplesf_full(2 * spec.VectorizedState, spec.maxActions, spec.FeatureSize, 128, 5, seed=1)
    buf = ag.PoolSample(spec, 200000)
    ag.mcts(ag.convert_back(net), 16, 1024, buf, spec=spec, cpuct=1.5, seed=3)
    L = buf.length_buffer()
    idx = np.arange(min(L, 2048))
    ev = (buf.state[idx], buf.policy[idx], buf.value[idx], buf.fstate[idx])
    tr = ag.Trainer.for_network(net, 2048)
    before = tr.lossTot(*ev)[0]
    tr.close()
    net, rep = ag.traininPipe(512, net, buf, epoch=2, verbose=False)
    tr = ag.Trainer.for_network(net, 2048)
    after = tr.lossTot(*ev)[0]
    tr.close()
    assert rep["batches"] == L // 512 - 1 and rep["samples_per_s"] > 0
    assert after < before, (before, after)
