"""GPU: randomly generated networks at the edges of what the ABI admits -- 1 gene, exactly 32 / 64 / 96 / 128
genes, 1..5 predictors per gene, arities 0..9 (constants, wide predictors), 1..8 action slots, wildcard targets --
against the oracle, which parses the same expression strings with its own evaluator."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODES = {"none": 0, "A": 1, "B": 2, "C": 3}


def _sop(names, table):
    """Sum-of-products expression string for a truth table over `names` (bit j of the row index = names[j])."""
    k = len(names)
    rows = [a for a in range(1 << k) if (table >> a) & 1]
    if k == 0 or not rows:
        return "( %s & ~ %s )" % (names[0] if names else "x0", names[0] if names else "x0") if not rows else "( x0 | ~ x0 )"
    if len(rows) == 1 << k:
        return "( %s | ~ %s )" % (names[0], names[0])
    terms = []
    for a in rows:
        terms.append("( " + " & ".join(("%s" if (a >> j) & 1 else "~ %s") % names[j] for j in range(k)) + " )")
    return " | ".join(terms)


def make_network(n, kmax, amax, seed, uniform=True):
    rng = np.random.default_rng(seed)
    genes = ["x%d" % i for i in range(n)]
    exprs = []
    for i in range(n):
        k = int(rng.integers(1, kmax + 1))
        row = []
        weights = np.ones(k) if uniform else rng.integers(1, 6, size=k).astype(float)
        for f in range(k):
            ar = int(rng.integers(0, min(amax, n) + 1))
            ins = sorted(rng.choice(n, size=ar, replace=False).tolist())
            table = int(rng.integers(0, 1 << min(1 << ar, 62))) | (int(rng.integers(0, 1 << 62)) << 62 if ar > 5 else 0)
            table &= (1 << (1 << ar)) - 1
            if ar >= 7:   # keep the expression strings of wide predictors short: few minterms
                table = 0
                for a in rng.integers(0, 1 << ar, size=5):
                    table |= 1 << int(a)
            row.append((_sop([genes[j] for j in ins], table), float(weights[f] / weights.sum())))
        exprs.append(row)
    return genes, exprs


def make_attractors(n, count, seed):
    rng = np.random.default_rng(seed)
    out = []
    for a in range(count):
        states = []
        for _ in range(int(rng.integers(1, 4))):
            s = [int(v) for v in rng.integers(0, 2, size=n)]
            if a % 2 and n > 2:
                for j in rng.choice(n, size=min(2, n), replace=False):
                    s[int(j)] = "*"
            states.append(tuple(s))
        out.append(states)
    return out


CASES = [
    # n, kmax, amax, bins, uniform, kernels
    (1, 3, 1, 1, True, ("sliced", "scalar")),
    (2, 4, 2, 3, True, ("sliced", "scalar")),
    (31, 3, 4, 3, True, ("sliced", "scalar")),
    (32, 4, 5, 8, True, ("sliced", "scalar")),
    (33, 2, 6, 2, True, ("sliced", "scalar")),
    (64, 3, 4, 3, True, ("sliced", "scalar")),
    (65, 3, 4, 3, True, ("sliced", "scalar")),
    (96, 2, 3, 4, True, ("sliced", "scalar")),
    (97, 3, 3, 3, True, ("scalar",)),           # more than 96 genes: general kernel only
    (128, 2, 4, 5, True, ("scalar",)),
    (20, 5, 3, 3, False, ("sliced", "scalar")),  # five predictors, real selection probabilities: third selection plane
    (26, 8, 4, 3, True, ("sliced", "scalar")),   # up to eight predictors, uniform
    (12, 9, 3, 3, True, ("scalar",)),            # more than eight predictors: general kernel only
    (24, 2, 9, 3, False, ("sliced", "scalar")),  # wide predictors (7..9 inputs), weighted selection
    (40, 3, 12, 3, True, ("sliced", "scalar")),  # up to 12 inputs, two-word states
    (20, 2, 14, 3, True, ("scalar",)),           # more than 12 inputs: general kernel only
]


@pytest.mark.parametrize("n,kmax,amax,bins,uniform,kernels", CASES)
@pytest.mark.parametrize("mode", ["A", "C"])
def test_synthetic_networks_bit_exact(n, kmax, amax, bins, uniform, kernels, mode):
    import torch
    from oracle import pbn_oracle as O
    from pbn_rl_b200 import AttractorSet, PBNNetwork, VecPBNEnv
    genes, exprs = make_network(n, kmax, amax, seed=1000 + n)
    net = PBNNetwork.from_expressions(genes, exprs)
    onet = O.OracleNetwork(genes, exprs)
    attrs = make_attractors(n, 5, seed=n)
    aset = AttractorSet(attrs, n)
    tables = O.attractor_tables(attrs, n)
    e, p, seed = 1500, 0.03, 4242
    rng = np.random.default_rng(n)
    w = net.n_words
    masks = np.array(net.state_mask(), dtype=np.uint64)
    state0 = ((rng.integers(0, 2**63, size=(e, w), dtype=np.int64).astype(np.uint64) * np.uint64(2))
              + rng.integers(0, 2, size=(e, w)).astype(np.uint64)) & masks
    target = rng.integers(-1, 5, size=e).astype(np.int32)
    kw = dict(horizon=3, r_success=2.5, r_step=-0.5, r_action=-1.0)
    for kernel in kernels:
        env = VecPBNEnv(net, e, aset, device="cuda:0", perturb_p=p, perturb_mode=mode, seed=seed, bins=bins, kernel=kernel, **kw)
        assert env.kernel == kernel
        env.set_state(torch.from_numpy(state0.astype(np.int64)), packed=True)
        env.set_target(torch.from_numpy(target))
        state, t = state0, np.zeros(e, np.uint16)
        ids = np.arange(e, dtype=np.uint64)
        for step in range(3):
            act = rng.integers(0, n + 1, size=(e, bins), dtype=np.uint8)
            if step == 1:      # injected randomness
                sel = np.stack([rng.integers(0, len(r), size=e) for r in exprs], axis=1).astype(np.uint8)
                pert = (rng.integers(0, 2**63, size=(e, w), dtype=np.int64).astype(np.uint64) & masks) * (rng.random((e, 1)) < 0.3)
                pert = pert.astype(np.uint64)
                env.step_injected(torch.from_numpy(act), torch.from_numpy(sel), torch.from_numpy(pert.astype(np.int64)))
                env.step_ctr = step + 1
            else:
                env.step(torch.from_numpy(act).cuda())
                if kernel == "scalar":
                    sel = O.scalar_stream_selection(onet, ids, step, seed)
                    pert = O.scalar_stream_perturbation(n, p, ids, step, seed)
                else:
                    sel, pert = O.sliced_stream(onet, p, ids, step, seed)
            nxt, t1, rew, term, trunc = O.batched_step(onet, tables, state, act, target, t, mode=MODES[mode], sel=sel, pert=pert, **kw)
            torch.cuda.synchronize()
            tag = "n=%d %s %s step %d" % (n, kernel, mode, step)
            assert np.array_equal(env.state.cpu().numpy().astype(np.uint64), nxt), tag
            assert np.array_equal(env.reward.cpu().numpy(), rew), tag
            assert np.array_equal(env.terminated.cpu().numpy(), term) and np.array_equal(env.truncated.cpu().numpy(), trunc), tag
            assert np.array_equal(env.t.cpu().numpy().astype(np.uint16), t1), tag
            state, t = nxt, t1
        if kernel == "sliced":      # uncontrolled rollout on the same odd-shaped network
            ref = env.state.clone()
            twin = VecPBNEnv(net, e, aset, device="cuda:0", perturb_p=p, perturb_mode=mode, seed=seed, bins=bins, kernel=kernel, **kw)
            twin.set_state(ref, packed=True)
            twin.step_ctr = env.step_ctr
            for _ in range(4):
                twin.step(None)
            env.rollout(4)
            assert torch.equal(env.state, twin.state), "rollout n=%d %s" % (n, mode)
            twin.close()
        env.close()
