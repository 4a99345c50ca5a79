"""Shared fixtures for the tests: golden networks, oracle twins, seeded inputs."""
import json
from functools import lru_cache
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"
NETS = ("pbn7", "pbn10", "pbn28", "pbn70")


@lru_cache(maxsize=None)
def product_net(name):
    from pbn_rl_b200 import PBNNetwork
    return PBNNetwork.from_json(GOLD / f"{name}.json")


@lru_cache(maxsize=None)
def oracle_net(name):
    from oracle.pbn_oracle import OracleNetwork
    return OracleNetwork.from_json(GOLD / f"{name}.json")


def golden(name):
    return json.loads((GOLD / name).read_text())


@lru_cache(maxsize=None)
def attractor_set(name):
    """Attractor table used with each golden network:
    pbn7 -> data/attractors_Bittner-7.pkl; pbn10 -> the 3 sink SCCs (K5);
    pbn28 -> data/attractors_Bittner-28.pkl permuted to file order (K3);
    pbn70 -> 16 synthetic single-state targets from a seeded uncontrolled oracle-free construction."""
    from pbn_rl_b200 import AttractorSet, sorted_id_permutation
    net = product_net(name)
    if name == "pbn7":
        return AttractorSet([[tuple(s) for s in a] for a in golden("attractors_bittner7.json")["attractors"]], 7)
    if name == "pbn10":
        sinks = golden("k5_stg.json")["pbn10"]["sink_sccs"]
        return AttractorSet([[tuple((s >> i) & 1 for i in range(10)) for s in m] for m in sinks], 10)
    if name == "pbn28":
        raw = AttractorSet([[tuple(s) for s in a] for a in golden("attractors_bittner28.json")["attractors"]], 28)
        return raw.permuted(sorted_id_permutation(net.genes))
    rng = np.random.default_rng(70)
    states = rng.integers(0, 2, size=(16, 70))
    return AttractorSet([[tuple(int(v) for v in s)] for s in states], 70)


def k4_inputs(n):
    mask_lo = (1 << min(n, 64)) - 1
    mask_hi = (1 << (n - 64)) - 1 if n > 64 else 0
    w = 1 if n <= 64 else 2
    out = np.zeros((4096, w), dtype=np.uint64)
    for j in range(4096):
        out[j, 0] = ((j + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF & mask_lo
        if w == 2:
            out[j, 1] = ((j + 1) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF & mask_hi
    return out


def k4_selections(n):
    """The four selection modes of fixture K4, in hash order."""
    j = np.arange(4096)[:, None]
    i = np.arange(n)[None, :]
    return [np.full((4096, n), k, dtype=np.uint8) for k in range(3)] + [((i + j) % 3).astype(np.uint8)]


def random_case(name, e, seed, pert_density=0.05, sel_max=3):
    """Seeded random inputs for one step of E envs."""
    net = product_net(name)
    n, w = net.n_genes, net.n_words
    rng = np.random.default_rng(seed)
    masks = np.array(net.state_mask(), dtype=np.uint64)
    state = rng.integers(0, 2**63, size=(e, w), dtype=np.int64).astype(np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(e, w)).astype(np.uint64)
    state &= masks
    actions = rng.integers(0, n + 1, size=(e, 3), dtype=np.uint8)
    sel = np.zeros((e, n), dtype=np.uint8)
    for i, fs in enumerate(net.functions):
        sel[:, i] = rng.integers(0, min(len(fs), sel_max), size=e)
    pert_bits = rng.random((e, n)) < pert_density
    pert = np.zeros((e, w), dtype=np.uint64)
    for i in range(n):
        pert[:, i >> 6] |= pert_bits[:, i].astype(np.uint64) << np.uint64(i & 63)
    keep = rng.random(e) < 0.5  # half the envs unperturbed, so PERT_A exercises both branches
    pert[keep] = 0
    n_attr = len(attractor_set(name))
    target = rng.integers(0, n_attr, size=e, dtype=np.int32)
    t = rng.integers(0, 25, size=e).astype(np.uint16)
    return dict(state=state, actions=actions, sel=sel, pert=pert, target=target, t=t)
