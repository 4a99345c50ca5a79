"""The C-ABI shared library without a GPU: it loads, exports every declared symbol, fails loudly on
compute calls, and its load-time code generator emits correct LOP3 trees (checked by compiling the
generated source as host C++)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import NETS, product_net
from pbn_rl_b200 import PBNNetwork, _cabi

ROOT = Path(__file__).resolve().parent.parent


def _lib():
    if not _cabi.LIB_PATH.exists():
        pytest.skip("libpbn_b200.so not built (run __graft_entry__.build())")
    return _cabi.load_library()


def test_library_exports_every_header_symbol():
    lib = _lib()
    header = (ROOT / "include" / "pbn_b200.h").read_text()
    declared = set(re.findall(r"\b(pbn_[a-z_0-9]+)\s*\(", header)) - {"pbn_step_args", "pbn_net_desc"}
    assert declared == set(_cabi.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.pbn_version()


def test_struct_layouts_match_the_header(tmp_path):
    """gcc compiles include/pbn_b200.h (it is plain C) and prints sizeof/offsetof of every struct and field;
    the ctypes mirrors in pbn_rl_b200/_cabi.py must agree field by field."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    structs = {"pbn_net_desc": _cabi.NetDesc, "pbn_step_args": _cabi.StepArgs, "pbn_host_io": _cabi.HostIO,
               "pbn_replay": _cabi.Replay}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "pbn_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", str(root / "include"), "-o", str(exe), str(src)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_no_silent_cpu_fallback():
    """Without a CUDA device pbn_create reports a CUDA error; nothing computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib()
    from pbn_rl_b200.vec_env import make_desc
    d, keep = make_desc(product_net("pbn7"))
    h = C.c_void_p()
    rc = lib.pbn_create(C.byref(d), C.byref(h))
    assert rc == -2 and not h.value
    assert b"cuda" in lib.pbn_last_error().lower()
    from pbn_rl_b200 import VecPBNEnv
    with pytest.raises(RuntimeError):
        VecPBNEnv(product_net("pbn7"), 16)
    del keep


def test_descriptor_validation_errors():
    lib = _lib()
    from pbn_rl_b200.vec_env import make_desc
    d, keep = make_desc(product_net("pbn7"))
    d.n_genes = 0
    h = C.c_void_p()
    assert lib.pbn_create(C.byref(d), C.byref(h)) == -1
    d, keep = make_desc(product_net("pbn7"), bins=99)
    assert lib.pbn_create(C.byref(d), C.byref(h)) == -1
    assert b"bins" in lib.pbn_last_error()
    assert lib.pbn_step(None, None, None) == -1
    del keep


def test_sliced_eligibility():
    from pbn_rl_b200.vec_env import jit_source
    assert "pbn_update_part" in jit_source(product_net("pbn28"))
    nonuniform = PBNNetwork.from_expressions(["a", "b"], [[("a | b", 0.9), ("a & b", 0.1)], ["a"]])
    assert "draw_weighted" in jit_source(nonuniform)   # arbitrary probabilities: threshold comparison, still bit-sliced
    five = PBNNetwork.from_expressions(["a", "b"], [["a", "b", "a|b", "a&b", "~a"], ["a"]])
    src = jit_source(five)                            # 5..8 predictors: a third selection plane
    assert "#define PBN_SELBITS 3" in src and "draw_weighted8" in src
    assert "#define PBN_SELBITS 2" in jit_source(nonuniform)
    nine = PBNNetwork.from_expressions(["a", "b"], [["a", "b", "a|b", "a&b", "~a", "~b", "a & ~b", "~a & b", "~a | b"], ["a"]])
    with pytest.raises(_cabi.PbnError):
        jit_source(nine)


HOST_HARNESS = r"""
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define __device__
#define __forceinline__ inline
#define __constant__ const
#define PBN_RNG_SELECT 0u
#define PBN_RNG_FIX 3u
static inline bool __any_sync(unsigned, bool p) { return p; }   // one column = one lane
#include "philox.cuh"
using pbn::Philox4;
using pbn::philox_stream_rk;
using pbn::draw_weighted;
using pbn::draw_weighted8;
template <int IMM> static inline uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) {
    const int idx = (((a >> i) & 1) << 2) | (((b >> i) & 1) << 1) | ((c >> i) & 1);
    r |= (uint32_t)((IMM >> idx) & 1) << i;
  }
  return r;
}
static inline uint32_t bmux(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (~a & c); }
#include "net_gen.cuh"
#include "net_update.inc"
int main(int argc, char** argv) {
  // stdin: N NSEL cases, then per case: N input planes, NSEL s0 planes, NSEL s1 planes
  // stdout per case: the out planes of pbn_update_part (row-format kernel), then those of pbn_eval_part
  // (plane-resident kernel) followed by its target-difference word;
  // then: D draws, each "gid step seed" -> the lo / hi selection planes of all slots from pbn_draw_part
  int n, nsel, cases;
  if (scanf("%d %d %d", &n, &nsel, &cases) != 3) return 1;
  const int nw = (n + 31) / 32;
  // (PBN_SELBITS == 3: a third selection plane per slot, read after s1 and laid out behind it)
  static uint32_t x[128 * 32], o[128 * 32], o2[128 * 32], tg[128 * 32], s0[128 * 32], s1[2 * 128 * 32];
  uint32_t* const s2 = s1 + 128 * 32;
  for (int c = 0; c < cases; ++c) {
    for (int i = 0; i < n; ++i) if (scanf("%u", &x[i * 32]) != 1) return 1;
    for (int i = 0; i < nsel; ++i) if (scanf("%u", &s0[i * 32]) != 1) return 1;
    for (int i = 0; i < nsel; ++i) if (scanf("%u", &s1[i * 32]) != 1) return 1;
#if PBN_SELBITS == 3
    for (int i = 0; i < nsel; ++i) if (scanf("%u", &s2[i * 32]) != 1) return 1;
#endif
    for (int i = 0; i < nw * 32; ++i) o[i * 32] = 0xDEADBEEFu;
#if PBN_SELBITS == 3
    for (uint32_t w = 0; w < 4; ++w) pbn::pbn_update_part(w, x, o, s0, s1, s2);
#else
    for (uint32_t w = 0; w < 4; ++w) pbn::pbn_update_part(w, x, o, s0, s1);
#endif
    for (int i = 0; i < nw * 32; ++i) printf("%u ", o[i * 32]);
    printf("\n");
    uint32_t d = 0;
    for (int i = 0; i < n; ++i) { o2[i * 32] = 0u; tg[i * 32] = x[((i + 1) % n) * 32]; }
    for (uint32_t q = 0; q < 8; ++q) {
      uint32_t lo[PBN_MAXS4], hi[PBN_MAXS4];   // the planes of group q mod 4
      for (int k = 0; k < PBN_MAXS4; ++k) { const int r = (int)(q & 3u) + 4 * k; lo[k] = r < nsel ? s0[r * 32] : 0u; hi[k] = r < nsel ? s1[r * 32] : 0u; }
#if PBN_SELBITS == 3
      uint32_t h2[PBN_MAXS4];
      for (int k = 0; k < PBN_MAXS4; ++k) { const int r = (int)(q & 3u) + 4 * k; h2[k] = r < nsel ? s2[r * 32] : 0u; }
      d |= pbn::pbn_eval_part<0>(q, x, o2, tg, 0u, lo, hi, h2);
#else
      d |= pbn::pbn_eval_part<0>(q, x, o2, tg, 0u, lo, hi);
#endif
    }
    for (int i = 0; i < n; ++i) printf("%u ", o2[i * 32]);
    printf("%u\n", d);
  }
  int draws;
  if (scanf("%d", &draws) != 1) return 1;
  for (int c = 0; c < draws; ++c) {
    unsigned long long gid, step, seed;
    if (scanf("%llu %llu %llu", &gid, &step, &seed) != 3) return 1;
    uint32_t rk[20];
    for (int r = 0; r < 10; ++r) { rk[2 * r] = (uint32_t)seed + r * 0x9E3779B9u; rk[2 * r + 1] = (uint32_t)(seed >> 32) + r * 0xBB67AE85u; }
    static uint32_t L[1024], H[1024], H2[1024];
    for (uint32_t q = 0; q < 4; ++q) {
      uint32_t lo[PBN_MAXS4], hi[PBN_MAXS4], h2[PBN_MAXS4] = {};
#if PBN_SELBITS == 3
      pbn::pbn_draw_group(q, gid, step, rk, lo, hi, h2);
#else
      pbn::pbn_draw_group(q, gid, step, rk, lo, hi);
#endif
      for (int k = 0; k < PBN_MAXS4; ++k) { const int r = (int)q + 4 * k; if (r < nsel) { L[r] = lo[k]; H[r] = hi[k]; H2[r] = h2[k]; } }
      // the split form of the 8-warp plane-resident kernel: even private blocks (PH = 0) and odd ones (PH = 1) drawn
      // apart, the odd results handed over, then the shared pool alone (PH = 2) -- must equal the one-call form
      uint32_t lo_e[PBN_MAXS4], hi_e[PBN_MAXS4], h2_e[PBN_MAXS4], lo_o[PBN_MAXS4], hi_o[PBN_MAXS4], h2_o[PBN_MAXS4];
      for (int k = 0; k < PBN_MAXS4; ++k) { lo_e[k] = hi_e[k] = h2_e[k] = 0xA5A5A5A5u; lo_o[k] = hi_o[k] = h2_o[k] = 0x5A5A5A5Au; }
#if PBN_SELBITS == 3
      pbn::pbn_draw_group<0>(q, gid, step, rk, lo_e, hi_e, h2_e);
      pbn::pbn_draw_group<1>(q, gid, step, rk, lo_o, hi_o, h2_o);
#else
      pbn::pbn_draw_group<0>(q, gid, step, rk, lo_e, hi_e);
      pbn::pbn_draw_group<1>(q, gid, step, rk, lo_o, hi_o);
      for (int k = 0; k < PBN_MAXS4; ++k) { h2_e[k] = 0u; h2_o[k] = 0u; }
#endif
      for (int k = 1; k < PBN_MAXS4; k += 2) { lo_e[k] = lo_o[k]; hi_e[k] = hi_o[k]; h2_e[k] = h2_o[k]; }
#if PBN_SELBITS == 3
      pbn::pbn_draw_group<2>(q, gid, step, rk, lo_e, hi_e, h2_e);
#else
      pbn::pbn_draw_group<2>(q, gid, step, rk, lo_e, hi_e);
#endif
      for (int k = 0; k < PBN_MAXS4; ++k)
        if (lo_e[k] != lo[k] || hi_e[k] != hi[k] || h2_e[k] != h2[k]) { fprintf(stderr, "split draw differs: group %u slot index %d\n", q, k); return 3; }
    }
    for (int r = 0; r < nsel; ++r) printf("%u %u %u ", L[r], H[r], H2[r]);
    printf("\n");
  }
  return 0;
}
"""


def _mixed_network():
    """K in {1,2,3,4}, uniform and weighted selection mixed, so that every branch of the generated draw runs."""
    genes = ["g%d" % i for i in range(11)]
    fs = [
        [("g1 | g2", 0.9), ("g1 & g2", 0.1)],                       # K=2 weighted
        ["g0"],                                                      # K=1
        ["g3 & g4", "g3 | g4", "~g5"],                               # K=3 uniform
        [("g0", 0.25), ("g1", 0.25), ("g2 & ~g3", 0.5)],              # K=3 weighted
        ["g4", "~g4"],                                               # K=2 uniform
        ["g1", "g2", "g3", "g4 & g0"],                               # K=4 uniform
        [("g6", 0.1), ("g7", 0.2), ("g8", 0.3), ("g9", 0.4)],        # K=4 weighted
        ["g7 | g8", "g9", "g10"],                                    # K=3 uniform
        ["g8"],                                                      # K=1
        [("g9 & g10", 0.999), ("g0", 0.001)],                        # K=2 weighted, extreme
        ["g10", "g5", "g6 & g7"],                                    # K=3 uniform
    ]
    return genes, fs


def _many_predictor_network():
    """Genes with 5..8 predictors (uniform and weighted) next to 1..4: three selection planes (PBN_SELBITS == 3)."""
    genes = ["m%d" % i for i in range(9)]
    fs = [
        ["m1", "m2", "m3", "m4", "m5"],                                                   # K=5 uniform
        [("m0", 0.05), ("m2", 0.1), ("m3", 0.15), ("m4", 0.2), ("m5", 0.2), ("m6 & m7", 0.3)],   # K=6 weighted
        ["m0 | m1", "m3", "m4", "m5", "m6", "m7", "~m8"],                                 # K=7 uniform
        [("m%d" % j, 0.125) for j in range(8)],                                           # K=8 uniform
        ["m0", "m1", "m2"],                                                               # K=3 uniform (pair-plane draw)
        ["m8"],                                                                           # K=1
        [("m1 & m2", 0.7), ("m3", 0.3)],                                                  # K=2 weighted
        ["m2", "m3 | m4", "m5", "m6"],                                                    # K=4 uniform
        [("m0", 0.01), ("m1", 0.01), ("m2", 0.01), ("m3", 0.01), ("m4", 0.01), ("m5", 0.01), ("m6", 0.01), ("m7", 0.93)],  # K=8 skewed
    ]
    return genes, fs


def _wide_network():
    """Predictors of 7..12 inputs (multi-word truth tables, Shannon-expanded LOP3 trees): rule-like sums of products,
    a parity (XOR decomposition), a dense random table, next to narrow ones."""
    rng = np.random.default_rng(77)
    genes = ["w%d" % i for i in range(14)]

    def sop(ins, minterms):
        return " | ".join("( " + " & ".join(("%s" if (a >> j) & 1 else "~ %s") % genes[ins[j]] for j in range(len(ins))) + " )"
                          for a in minterms)

    fs = []
    for i in range(14):
        ar = [7, 8, 9, 10, 11, 12, 3, 7, 12, 2, 8, 9, 10, 1][i]
        ins = sorted(rng.choice(14, size=ar, replace=False).tolist())
        if i == 7:     # parity of 7 inputs
            fs.append([sop(ins, [a for a in range(128) if bin(a).count("1") & 1])])
        elif i == 10:  # dense random table of 8 inputs
            fs.append([sop(ins, [a for a in range(256) if rng.integers(0, 2)])])
        elif i == 11:  # two wide predictors under one selection slot
            fs.append([sop(ins, rng.integers(0, 1 << ar, size=6).tolist()), " & ".join(genes[j] for j in ins[:8])])
        elif i == 12:  # an OR of 10 inputs, an AND of 3, a wide sum of products
            fs.append([" | ".join(genes[j] for j in ins), " & ".join(genes[j] for j in ins[:3]),
                       sop(ins, rng.integers(0, 1 << ar, size=4).tolist())])
        else:
            fs.append([sop(ins, rng.integers(0, 1 << ar, size=7).tolist())])
    return genes, fs


@pytest.mark.parametrize("name", NETS + ("mixed", "wide", "many"))
def test_generated_lop3_trees_match_truth_tables(name, tmp_path):
    """Compile the generated net_gen.cuh / net_update.inc with g++: the predictor trees of both kernels against the
    truth tables on random bit-planes, and the generated selection draw (pbn_draw_group, with csrc/philox.cuh compiled
    for the host) against the oracle's twin of the stream (oracle/pbn_oracle.py: sliced_stream)."""
    _lib()
    from oracle import pbn_oracle as O
    from helpers import oracle_net
    from pbn_rl_b200.vec_env import jit_source
    if name in ("mixed", "wide", "many"):
        genes, fs = {"mixed": _mixed_network, "wide": _wide_network, "many": _many_predictor_network}[name]()
        net = PBNNetwork.from_expressions(genes, fs)
        onet = O.OracleNetwork(genes, [[(f, 1.0 / len(g)) if isinstance(f, str) else f for f in g] for g in fs])
    else:
        net, onet = product_net(name), oracle_net(name)
    src = jit_source(net)
    gen, upd = src.split("// ---- net_gen.cuh\n")[1].split("// ---- net_update.inc\n")
    (tmp_path / "net_gen.cuh").write_text(gen)
    (tmp_path / "net_update.inc").write_text(upd)
    (tmp_path / "h.cpp").write_text(HOST_HARNESS)
    exe = tmp_path / "h"
    csrc = ROOT / "pbn_rl_b200" / "csrc"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(csrc), "-o", str(exe), str(tmp_path / "h.cpp")], check=True)
    n = net.n_genes
    slots = [i for i, fs in enumerate(net.functions) if len(fs) > 1]
    three = max(len(fs) for fs in net.functions) > 4   # the network carries a third selection plane
    rng = np.random.default_rng(1)
    cases = 6
    lines = ["%d %d %d" % (n, len(slots), cases)]
    data = []
    for _ in range(cases):
        x = rng.integers(0, 2**32, size=n, dtype=np.uint64)
        sel = np.stack([rng.integers(0, len(net.functions[i]), size=32) for i in slots], axis=0) if slots else np.zeros((0, 32), int)
        s0 = [(int(sum(int(v & 1) << b for b, v in enumerate(row)))) for row in sel]
        s1 = [(int(sum(int((v >> 1) & 1) << b for b, v in enumerate(row)))) for row in sel]
        s2 = [(int(sum(int((v >> 2) & 1) << b for b, v in enumerate(row)))) for row in sel]
        lines.append(" ".join(str(int(v)) for v in x) + " " + " ".join(map(str, s0)) + " " + " ".join(map(str, s1))
                     + ((" " + " ".join(map(str, s2))) if three else ""))
        data.append((x, sel))
    draws = [(5, 0, 0x5EED), (1234567, 3, 0x5EED), ((1 << 33) + 9, (1 << 40) + 7, 0xDEADBEEF12345678)]
    lines.append(str(len(draws)))
    lines += ["%d %d %d" % d for d in draws]
    out = subprocess.run([str(exe)], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    rows = [[int(v) for v in line.split()] for line in out.strip().splitlines()]
    for k, (x, sel) in enumerate(data):
        got, got2 = rows[2 * k], rows[2 * k + 1]
        for b in range(32):
            state = sum(((int(x[i]) >> b) & 1) << i for i in range(n))
            choice = [0] * n
            for j, i in enumerate(slots):
                choice[i] = int(sel[j][b])
            want = net.next_state_int(state, choice)
            assert sum(((got[i] >> b) & 1) << i for i in range(n)) == want
            assert sum(((got2[i] >> b) & 1) << i for i in range(n)) == want
        for i in range(n, len(got)):
            assert got[i] == 0  # unused planes are cleared
        diff = 0
        for i in range(n):
            diff |= got2[i] ^ int(x[(i + 1) % n])
        assert got2[n] == diff  # OR of (next state XOR target planes)
    for (gid, step, seed), got in zip(draws, rows[2 * cases:]):
        tile, lane = gid >> 5, gid & 31
        ids = np.array([tile * 1024 + 128 * (b >> 2) + 4 * lane + (b & 3) for b in range(32)], dtype=np.uint64)
        sel, _ = O.sliced_stream(onet, 0.0, ids, step, seed)
        for r, i in enumerate(slots):
            lo, hi, h2 = got[3 * r], got[3 * r + 1], got[3 * r + 2]
            have = [((lo >> b) & 1) + 2 * ((hi >> b) & 1) + 4 * ((h2 >> b) & 1) for b in range(32)]
            assert have == [int(v) for v in sel[:, i]], (name, gid, step, r)
