"""The PyTorch DQN restatement: architecture facts on CPU, the batched agent on GPU."""
import numpy as np
import pytest
import torch

import oracle


def test_model_has_the_reference_parameter_count():
    """Dqn8TestNOPERCNN.py:17 '198 milioni di parametri'; exact Keras count 197,204,996."""
    from g2048 import dqn
    with torch.device("meta"):
        m = dqn.DQNModel()
    assert sum(p.numel() for p in m.parameters()) == 197_204_996


def test_forward_shapes_and_same_padding_on_cpu():
    from g2048 import dqn
    torch.manual_seed(0)
    m = dqn.DQNModel(width=16, hidden=8).eval()
    boards = np.array([0x0000000000000021, 0x123456789ABCDEF1], np.uint64)
    x = torch.from_numpy(oracle.encode_onehot(boards))
    y = m(x)
    assert y.shape == (2, 4) and torch.isfinite(y).all()
    # 'same' for the even kernels pads (k-1)//2 before and k//2 after, like TensorFlow
    conv = m.blocks[0].convs[1]           # kernel 2
    xin = x.permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(xin, (0, 1, 0, 1)), conv.weight, conv.bias)
    assert torch.allclose(conv(xin), ref, atol=1e-6)


@pytest.mark.gpu
def test_batched_dqn_agent_trains_on_gpu_envs():
    import g2048
    from g2048 import dqn
    torch.manual_seed(0)
    n = 4096
    env = g2048.BatchedGame2048Env(n, "nopenalty", seed=11)
    agent = dqn.BatchedDQNAgent(width=32, hidden=64, memory_size=1 << 16, batch_size=256, epsilon=0.5, learning_rate=1e-3)
    env.reset()
    # encode_state == the oracle's restatement of Dqn8TestNOPERCNN.py:271-277
    enc = agent.encode_state(env.boards).cpu().numpy()
    assert np.array_equal(enc, oracle.encode_onehot(env.boards.cpu().numpy().view(np.uint64)))
    for t in range(40):
        state = env.boards.clone()
        legal = env.legal_mask()
        a = agent.act_ripetitive(state, legal)
        ok = ((legal.to(torch.int64) >> a.to(torch.int64)) & 1).bool() | (legal == 0)
        assert bool(ok.all())                      # never an illegal move while a legal one exists
        dqn.dqn_step(env, agent)
    assert agent.nb_entries == 40 * n and agent.nb_entries <= agent.memory_size or agent.nb_entries == agent.memory_size
    losses = [agent.replay() for _ in range(60)]
    assert all(np.isfinite(l) for l in losses)
    assert np.mean(losses[-10:]) < np.mean(losses[:10])
    agent.update_target_model()
    for p, q in zip(agent.model.parameters(), agent.target_model.parameters()):
        assert torch.equal(p, q)
    # terminal bonus rule of the driver (mainDQL_CNN_step2.py:202-213)
    b = torch.tensor([0xB, 0xAA, 0xA9, 0x9], dtype=torch.int64, device="cuda")
    d = torch.tensor([True, True, True, True], device="cuda")
    assert dqn.terminal_bonus(b, d).tolist() == [100.0, 50.0, 0.0, 0.0]
