"""CPU: the replay restatement (oracle/replay_oracle.py) behaves like the reference's python-list memory."""
import numpy as np

from oracle import replay_oracle as R


def test_ring_overwrites_oldest_first():
    mem = R.OracleReplay(3)
    for k in range(5):
        mem.store(R.Transition((k,), (0,), np.array([k, 0, 0]), float(k), (k + 1,), k == 4))
    assert len(mem) == 3 and mem.current_index == 2
    assert [t.reward for t in mem.buffer] == [3.0, 4.0, 2.0]
    out = mem.batch_tensors([1, 2])
    assert out["obs"].shape == (2, 2, 1) and out["actions"].shape == (2, 3, 1) and out["done"].tolist() == [[1.0], [0.0]]
    assert out["obs"].dtype == np.float32 and out["actions"].dtype == np.int64


def test_bits_and_targets():
    assert R.words_to_bits([0b1011], 5) == (1, 1, 0, 1, 0)
    assert R.words_to_bits([0, 0b10], 66) == tuple([0] * 65 + [1])
    attrs = [[(1, "*", 0)], [(0, 1, 1), (1, 1, 1)]]
    assert R.target_state(attrs, 0, 3) == (1, 0, 0) and R.target_state(attrs, 1, 3) == (0, 1, 1)
    assert R.target_state(attrs, -1, 3) == (0, 0, 0)
    obs = R.observation(np.array([[0b101]], dtype=np.uint64), np.array([1]), attrs, 3)
    assert obs.tolist() == [[[1.0, 0.0, 1.0]], [[0.0, 1.0, 1.0]]]
