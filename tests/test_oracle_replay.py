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


def test_oracle_matches_reference_memory_fixture():
    """tests/golden/replay_expected.json was produced by running the reference's own ExperienceReplay
    (bdq_model/memory.py) and the tensor-building lines of update_policy (bdq_model/__init__.py:100-111):
    the restatement must reproduce the buffer after the scripted stores and every sampled batch."""
    import json
    from pathlib import Path
    fx = json.loads((Path(__file__).resolve().parent / "golden" / "replay_expected.json").read_text())
    mem = R.OracleReplay(fx["capacity"])
    for t in fx["script"]:
        mem.store(R.Transition(tuple(t["state"]), tuple(t["target"]), np.array(t["action"]), t["reward"],
                               tuple(t["next_state"]), t["done"]))
    assert len(mem) == fx["len"] and mem.current_index == fx["current_index"]
    for have, want in zip(mem.buffer, fx["buffer_after"]):
        assert list(have.state) == want["state"] and list(have.target) == want["target"]
        assert have.action.tolist() == want["action"] and have.reward == want["reward"]
        assert list(have.next_state) == want["next_state"] and bool(have.done) == want["done"]
    for smp in fx["samples"]:
        out = mem.batch_tensors(smp["index"])
        assert out["obs"].shape == tuple(smp["input_tuples_shape"]) and out["actions"].shape == tuple(smp["actions_shape"])
        assert out["obs"][0].tolist() == smp["states"] and out["obs"][1].tolist() == smp["targets"]
        assert out["next_obs"][0].tolist() == smp["next_states"] and out["next_obs"][1].tolist() == smp["targets"]
        assert out["actions"].tolist() == smp["actions"]
        assert out["reward"].tolist() == smp["rewards"] and out["done"].tolist() == smp["masks"]
    st, tg = fx["script"][0]["state"], fx["script"][0]["target"]
    bits = lambda v: sum(int(b) << i for i, b in enumerate(v))
    attrs = [[tuple(tg)]]
    obs = R.observation(np.array([[bits(st)]], dtype=np.uint64), np.array([0]), attrs, fx["n"])
    assert obs[:, 0, :].tolist() == fx["predict_input"]
