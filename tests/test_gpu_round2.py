"""GPU tests added in round 2 (run with -m gpu on the B200 box): the benchmark's own code path against the oracle at
BASELINE.json's full configuration, the device RNG as a distribution, the precision selection of the tower, the
device replay ring -> SGD batch path against the reference's train step, the cached ply graph's life cycle, and the
strength ordering of the shipped checkpoints.  Everything goes through the C ABI; the oracle is only the checker."""
import collections
import ctypes as C
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from harness import diff_tree, oracle_for
from oracle.mcts import OracleMCTS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


# --------------------------------------------------------------------------- full configuration vs the oracle
class RecordedOracle(OracleMCTS):
    """OracleMCTS that consumes the Dirichlet vectors the engine's Philox sampler produced (one per descent) and the
    (priors, value) rows the tensor-core tower produced for the very leaves it asks about."""

    def __init__(self, game):
        super().__init__(game, dirichlet=self._draw)
        self.noise, self.j, self.rows = None, 0, {}

    def _draw(self, alpha):
        z = self.noise[self.j]
        self.j += 1
        return z

    def evaluate(self, states, players, net, device="cpu"):
        pri = np.stack([self.rows[s][0] for s in states]).astype(np.float32)
        val = np.array([self.rows[s][1] for s in states], dtype=np.float32)
        return pri, val

    def minibatch(self, batch, state, player, noise_bj, rows):
        self.noise, self.j, self.rows = noise_bj, 0, rows
        self.search_minibatch(batch, state, player, None)


def test_full_configuration_pipeline_matches_oracle(torch_cuda):
    """BASELINE.json configs[1] through the code path bench.py times: 4,096 games (two pipeline parts of 2,048),
    search_batch(100, 8), Philox Dirichlet noise, the real bf16 tcgen05 tower, `play_multi` (one CUDA graph per ply), two
    plies.  A twin set of engines with the same seeds is stepped launch by launch through the C ABI, recording for eight
    sampled games the noise actually used and the tower's rows for their leaves; (1) the pipelined engines must equal
    the twins bit for bit (every node record, count, root), (2) the sampled games, replayed in the oracle with exactly
    those numbers, must give bit-identical N / W / Q / P trees, policies and moves."""
    torch = torch_cuda
    from caro_ai_b200 import _cabi
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.model import DeviceNet, Net
    game = ConnectFour()
    og = oracle_for(game)
    A, Bt, Cn, plies, parts, Gp, cap = 7, 8, 100, 2, 2, 2048, 2048
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")

    def engines():
        return [SelfPlayEngine(game, Gp, max_batch=Bt, node_capacity=cap, seed=500 + h) for h in range(parts)]

    fast = engines()
    SelfPlayEngine.play_multi(fast, dn, moves=plies, count=Cn, batch=Bt, tau_plies=10, auto_restart=True)
    twins = engines()
    sample = [0, 777, 1234, 2047]
    trees = {(h, g): RecordedOracle(og) for h in range(parts) for g in sample}
    lib = _cabi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    noise_buf = torch.empty((Gp, Bt, A), dtype=torch.float64, device="cuda")
    probs = torch.empty((Gp * Bt, A), dtype=torch.float32, device="cuda")
    values = torch.empty((Gp * Bt,), dtype=torch.float32, device="cuda")
    gidx = torch.tensor(sample, device="cuda")
    leaf_evals = 0
    for ply in range(plies):
        roots = {}
        for h, e in enumerate(twins):
            st, pl = e.roots()
            for g in sample:
                roots[(h, g)] = (st[g], pl[g])
        for mb in range(Cn):
            for h, e in enumerate(twins):
                e.select(Bt, mb, None, noise_out=noise_buf)
                e.plan(Bt)
                _cabi.check(lib.caro_net_forward(dn.handle, game.game_kind, 0, 0, e.region("leaf_board").data_ptr(),
                                                 e.region("leaf_player").data_ptr(), e.region("leaf_count").data_ptr(), Gp * Bt,
                                                 probs.data_ptr(), values.data_ptr(), dn.impl, stream))
                kind = e.region("desc_kind")[gidx].cpu().numpy()
                slot = e.region("desc_slot")[gidx].cpu().numpy()
                boards = e.region("desc_board")[gidx].cpu().numpy().view(np.uint64)
                nz = noise_buf[gidx].cpu().numpy()
                sl = torch.from_numpy(np.maximum(slot, 0).reshape(-1).astype(np.int64)).cuda()
                prow = probs[sl].cpu().numpy().reshape(len(sample), Bt, A)
                vrow = values[sl].cpu().numpy().reshape(len(sample), Bt)
                e.expand_backup(Bt, probs, values)
                for i, g in enumerate(sample):
                    rows = {}
                    states = game.states_from_boards(boards[i])
                    for j in range(Bt):
                        if kind[i, j] == 2 and slot[i, j] >= 0:
                            rows[states[j]] = (prow[i, j], vrow[i, j])
                            leaf_evals += 1
                    s, p = roots[(h, g)]
                    trees[(h, g)].minibatch(Bt, s, p, nz[i], rows)
        for h, e in enumerate(twins):
            pi_d, _, _ = e.root_policy(2, 10)
            pi_d = pi_d[gidx].cpu().numpy()
            actions = e.advance(10, None, auto_restart=True).cpu().numpy()
            st, pl = e.roots()
            for i, g in enumerate(sample):
                s, p = roots[(h, g)]
                t = trees[(h, g)]
                errs = diff_tree(e.export_tree(g), t, A)
                assert not errs, "part %d game %d ply %d: %s" % (h, g, ply, errs[:5])
                pi, _ = t.get_policy_value(s, tau=1)
                assert [float(x) for x in pi] == pi_d[i].tolist()
                assert pi[actions[g]] > 0
                s2, won = og.move(s, int(actions[g]), p)
                assert not won and st[g] == s2 and pl[g] == 1 - p
    assert leaf_evals > len(trees) * plies * 150
    torch.cuda.synchronize()
    for a, b in zip(fast, twins):
        ca, cb = a.counters(), b.counters()
        assert ca == cb and ca["errors"] == 0 and ca["descents"] == Gp * Bt * Cn * plies
        assert a.roots() == b.roots()
        for name in ("nodes", "node_count", "ply"):
            assert torch.equal(a.region(name), b.region(name)), name
    for e in fast + twins:
        e.close()
    dn.close()


# --------------------------------------------------------------------------- device RNG as a distribution
def test_device_dirichlet_noise_is_the_host_stream_and_a_dirichlet(torch_cuda):
    """noise_kernel (Philox 4x32-10 + Marsaglia-Tsang + boost) on the GPU: (1) every vector sums to one, (2) the draws are
    the addressed stream -- the host compile of the SAME rng.cuh reproduces them (the device uses fast-math log / exp / cos,
    so 'the same' is 1e-4 relative on >= 99.5 % of the components and never a different rejection path for most), (3)
    over 10^5 vectors the marginals have Dirichlet(0.3 x 7) mean 1/7 and variance (1/7)(6/7)/(7 x 0.3 + 1), the pairwise
    covariance is -(1/49)/(3.1), and different minibatch indices / games give different vectors."""
    torch = torch_cuda
    import subprocess
    from conftest import ROOT
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour
    game = ConnectFour()
    G, Bt, A, seed = 12800, 8, 7, 0x1234567887654321
    eng = SelfPlayEngine(game, G, max_batch=Bt, node_capacity=16, seed=seed)
    out = torch.empty((G, Bt, A), dtype=torch.float64, device="cuda")
    eng.select(Bt, 5, None, noise_out=out)
    z = out.cpu().numpy()
    out2 = torch.empty_like(out)
    eng.select(Bt, 6, None, noise_out=out2)
    z2 = out2.cpu().numpy()
    assert np.isfinite(z).all() and (z >= 0).all()
    np.testing.assert_allclose(z.sum(axis=2), 1.0, atol=1e-12)
    assert np.abs(z - z2).max() > 0.5 and not np.array_equal(z[0], z[1])
    flat = z.reshape(-1, A)
    a0 = A * 0.3
    np.testing.assert_allclose(flat.mean(axis=0), 1 / A, atol=3e-3)
    np.testing.assert_allclose(flat.var(axis=0), (1 / A) * (1 - 1 / A) / (a0 + 1), rtol=0.03)
    cov = np.cov(flat.T)
    off = cov[~np.eye(A, dtype=bool)]
    np.testing.assert_allclose(off, -(1 / A) ** 2 / (a0 + 1), rtol=0.1)
    # the heavy lower tail of Gamma(0.3): P(component < 1e-3) of a Dirichlet(0.3 x 7) marginal = Beta(0.3, 1.8) cdf
    from scipy import stats
    for q in (1e-3, 0.05, 0.5):
        assert abs((flat[:, 0] < q).mean() - stats.beta.cdf(q, 0.3, a0 - 0.3)) < 6e-3
    # host stream: same addresses (seed, uid, ply, side to move, minibatch * batch + descent, action)
    so = os.path.join(ROOT, "tests", "native", "libhostcheck.so")
    if not os.path.exists(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                               os.path.join(ROOT, "tests", "native", "hostcheck.cpp"), "-o", so])
    hc = C.CDLL(so)
    hc.hc_gamma.restype = C.c_float
    hc.hc_gamma.argtypes = [C.c_float] + [C.c_uint32] * 5
    uid = eng.region("uid").cpu().numpy().astype(np.uint64)
    ply = eng.region("ply").cpu().numpy()
    who = eng.region("root_player").cpu().numpy()
    k0 = (seed & 0xFFFFFFFF) ^ 0x44495243
    k1 = seed >> 32
    close = total = 0
    for g in range(0, G, 97):
        for j in range(Bt):
            gm = np.array([hc.hc_gamma(0.3, k0, k1, int(uid[g]) & 0xFFFFFFFF, ((int(uid[g]) >> 32) ^ (int(ply[g]) << 16) ^ int(who[g])) & 0xFFFFFFFF,
                                       ((5 * Bt + j) << 8) | a) for a in range(A)], dtype=np.float32)
            want = gm.astype(np.float64) / gm.astype(np.float64).sum()
            close += int(np.sum(np.abs(z[g, j] - want) <= 1e-4 * np.maximum(want, 1e-6) + 1e-9))
            total += A
    assert close >= 0.995 * total, (close, total)
    eng.close()


# --------------------------------------------------------------------------- precision selection
def test_auto_precision_keeps_the_contract_on_every_network(torch_cuda):
    """`DeviceNet(precision="auto")` -- what train.py, evaluate(), play.py, Session and the MCTS facade construct --
    against PyTorch fp32 on every test network, 1e-3 on priors AND values with no loosened gate: random-init networks
    stay on the one-pass bf16 tower, the shipped trained checkpoints (policy logits of +-100) are moved to the split
    mode by the on-device check, and `update()` re-runs the check (random -> trained -> random)."""
    import torch
    from test_gpu_parity import _net_cases, _reference_outputs
    from harness import random_position
    from caro_ai_b200.model import DeviceNet
    rng = np.random.default_rng(3)
    picked = {}
    nets = {}
    for tag, game, net in _net_cases():
        og = oracle_for(game)
        cells = game.obs_shape[1] * game.obs_shape[2]
        count = 300 if cells < 100 else 40
        pos = [random_position(og, rng, int(rng.integers(0, min(40, max(1, cells - 4))))) for _ in range(count)]
        states, players = [p[0] for p in pos], [p[1] for p in pos]
        ref_p, ref_v = _reference_outputs(game, net, states, players)
        dn = DeviceNet(net, game)
        assert dn.requested == "auto" and dn.calibration["positions"] == 256
        p, v = dn.forward_states(states, players)
        p, v = p.cpu().numpy(), v.cpu().numpy()
        dp, dv = np.abs(p - ref_p).max(), np.abs(v - ref_v).max()
        assert dp < 1e-3 and dv < 1e-3 and (p.argmax(1) == ref_p.argmax(1)).all(), (tag, dn.precision, dp, dv, dn.calibration)
        picked[tag] = dn.precision
        nets[tag] = (game, net, dn)
    # fastest one-pass tower that passes: fp16 on the row-tiled towers' boards, bf16 on large ones
    assert picked["c4-random"] == "fp16" and picked["mnk54-random"] == "fp16" and picked["caro-random"] == "bf16", picked
    cal = nets["c4-random"][2].calibration
    assert cal["fp16_max_abs_prior_diff"] < cal["bf16_max_abs_prior_diff"] < 1e-3 and cal["fp16_max_abs_value_diff"] < cal["bf16_max_abs_value_diff"]
    assert picked["c4-trained"] == "bf16x3", (picked, nets["c4-trained"][2].calibration)
    # update(): the same handle follows the weights it is given
    dn = nets["c4-random"][2]
    dn.update(nets["c4-trained"][1])
    assert dn.precision == "bf16x3" and dn.calibration["max_abs_prior_diff"] > 1e-3
    dn.update(nets["c4-random"][1])
    assert dn.precision == "fp16"
    for _, _, d in nets.values():
        d.close()


def test_split_precision_row_tiled_tower(torch_cuda):
    """net_rx.cu (impl 2 for boards <= 6 x 7: fp16 hi + lo activations and weights, three MMAs per product, block-major
    weight streaming) vs PyTorch fp32: 5e-4 on priors and values for the shipped trained checkpoints (policy logits of
    +-100; the one-pass bf16 tower is off by 0.2 there) and for random-init networks on every geometry class the
    row-tiled towers serve, ragged leaf counts and several passes per CTA included; bit-identical from run to run; and
    agreement with the tap-per-MMA split kernel (impl 4)."""
    torch = torch_cuda
    from test_gpu_parity import _net_cases, _random_net, _reference_outputs
    from harness import random_position
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import DeviceNet
    rng = np.random.default_rng(21)
    cases = [(tag, game, net, (300,)) for tag, game, net in _net_cases() if game.obs_shape[1] <= 6]
    cases += [("c4-ragged", ConnectFour(), _random_net(ConnectFour()), (1, 15, 16, 17, 2400, 5000)),
              ("ttt-ragged", TicTacToe(3, 3), _random_net(TicTacToe(3, 3)), (1, 31, 33, 4800)),
              ("mnk43", TicTacToe(4, 3), _random_net(TicTacToe(4, 3)), (5, 40)),
              ("mnk64", TicTacToe(6, 4), _random_net(TicTacToe(6, 4)), (3, 37))]
    for tag, game, net, counts in cases:
        og = oracle_for(game)
        cells = game.obs_shape[1] * game.obs_shape[2]
        dn = DeviceNet(net, game, precision="bf16x3")
        for count in counts:
            base = [random_position(og, rng, int(rng.integers(0, max(1, cells - 3)))) for _ in range(min(count, 300))]
            pos = [base[i % len(base)] for i in range(count)]
            states, players = [p[0] for p in pos], [p[1] for p in pos]
            ref_p, ref_v = _reference_outputs(game, net, states[:len(base)], players[:len(base)])
            p2, v2 = dn.forward_states(states, players, impl=2)
            p2b, v2b = dn.forward_states(states, players, impl=2)
            p4, v4 = dn.forward_states(states, players, impl=4)
            torch.cuda.synchronize()
            assert torch.equal(p2, p2b) and torch.equal(v2, v2b), (tag, count)
            p2, v2, p4, v4 = p2.cpu().numpy(), v2.cpu().numpy(), p4.cpu().numpy(), v4.cpu().numpy()
            assert np.isfinite(p2).all() and np.isfinite(v2).all()
            n = len(base)
            for i in range(0, count, n):
                m = min(n, count - i)
                dp, dv = np.abs(p2[i:i + m] - ref_p[:m]).max(), np.abs(v2[i:i + m] - ref_v[:m]).max()
                assert dp < 5e-4 and dv < 5e-4, (tag, count, i, dp, dv)
                assert (p2[i:i + m].argmax(1) == ref_p[:m].argmax(1)).all()
            assert np.abs(p2 - p4).max() < 1e-3 and np.abs(v2 - v4).max() < 1e-3
        dn.close()


def test_cta_pair_tower_is_bit_identical(torch_cuda):
    """impl 5 = the row-tiled bf16 tower run by clusters of two CTAs (tcgen05 cta_group::2: one MMA stream of M = 256, each
    CTA stores and fetches half of every B operand, accumulators zeroed by the epilogues instead of overwritten by the first
    MMA): the same products summed in the same order as impl 0, so priors and values must be BIT-identical -- for every
    geometry class (Connect4, pitch-4 3x3, 4x3-in-a-row, 6x6), odd and ragged group counts (a peer CTA with an empty group),
    several passes per pair, and with an SM limit (an odd one leaves one SM out of the pairs)."""
    torch = torch_cuda
    from test_gpu_parity import _random_net
    from harness import random_position
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import DeviceNet
    rng = np.random.default_rng(33)
    cases = [(ConnectFour(), (2, 17, 33, 47, 2400, 9473)), (TicTacToe(3, 3), (33, 65, 4800)), (TicTacToe(4, 3), (40,)),
             (TicTacToe(6, 4), (37, 700))]
    for game, counts in cases:
        og = oracle_for(game)
        cells = game.obs_shape[1] * game.obs_shape[2]
        dn = DeviceNet(_random_net(game), game, precision="bf16")
        base = [random_position(og, rng, int(rng.integers(0, max(1, cells - 3)))) for _ in range(200)]
        for count in counts:
            pos = [base[i % len(base)] for i in range(count)]
            states, players = [p[0] for p in pos], [p[1] for p in pos]
            p0, v0 = dn.forward_states(states, players, impl=0)
            p5, v5 = dn.forward_states(states, players, impl=5)
            torch.cuda.synchronize()
            assert torch.equal(p0, p5) and torch.equal(v0, v5), (game.obs_shape, count)
        for limit in (6, 7, 0):
            dn.set_grid_limit(limit)
            pos = [base[i % len(base)] for i in range(1500)]
            p0, v0 = dn.forward_states([p[0] for p in pos], [p[1] for p in pos], impl=0)
            p5, v5 = dn.forward_states([p[0] for p in pos], [p[1] for p in pos], impl=5)
            torch.cuda.synchronize()
            assert torch.equal(p0, p5) and torch.equal(v0, v5), (game.obs_shape, "limit", limit)
        dn.close()
    # the tap-per-MMA tower (large boards) has the same switch: impl 6 = impl 3 as CTA pairs, each CTA storing 32 of the 64
    # output channels of every tap
    for game, counts in ((TicTacToe(15, 5), (1, 3, 300, 1501)), (TicTacToe(9, 5), (7, 2000)), (ConnectFour(), (17, 1000))):
        og = oracle_for(game)
        dn = DeviceNet(_random_net(game), game, precision="bf16")
        base = [random_position(og, rng, int(rng.integers(0, 14))) for _ in range(100)]
        for count in counts:
            pos = [base[i % len(base)] for i in range(count)]
            states, players = [p[0] for p in pos], [p[1] for p in pos]
            p3, v3 = dn.forward_states(states, players, impl=3)
            p6, v6 = dn.forward_states(states, players, impl=6)
            torch.cuda.synchronize()
            assert torch.equal(p3, p6) and torch.equal(v3, v6), (game.obs_shape, count)
        dn.close()


def test_caro_heads_on_tensor_cores(torch_cuda):
    """Boards larger than 8 x 8 (Caro 15 x 15, 9 x 9) evaluate their FC heads as one split-precision tcgen05 GEMM per 128
    leaves (net_heads.cu) from features the tower exports: priors and values vs PyTorch fp32 at 1e-3 and vs the fp32 SIMT
    tower at 2e-4 for leaf counts around the 128-leaf tile and 2-board group boundaries, bit-identical from run to run,
    rows beyond a device-side count untouched."""
    torch = torch_cuda
    from test_gpu_parity import _random_net, _reference_outputs
    from harness import random_position
    from caro_ai_b200 import _cabi
    from caro_ai_b200.game import TicTacToe
    from caro_ai_b200.model import DeviceNet
    rng = np.random.default_rng(33)
    for n, k, counts in ((15, 5, (1, 2, 127, 128, 129, 300)), (9, 4, (3, 130))):
        game = TicTacToe(n, k)
        og = oracle_for(game)
        net = _random_net(game, seed=n)
        dn = DeviceNet(net, game, precision="bf16")
        base = [random_position(og, rng, int(rng.integers(0, 40 if n == 15 else 14))) for _ in range(48)]  # few enough plies for random play-outs to stay win-free
        ref_p, ref_v = _reference_outputs(game, net, [p[0] for p in base], [p[1] for p in base])
        for count in counts:
            pos = [base[i % len(base)] for i in range(count)]
            states, players = [p[0] for p in pos], [p[1] for p in pos]
            p0, v0 = dn.forward_states(states, players, impl=0)
            p0b, v0b = dn.forward_states(states, players, impl=0)
            p1, v1 = dn.forward_states(states, players, impl=1)
            p2, v2 = dn.forward_states(states, players, impl=2)
            torch.cuda.synchronize()
            assert torch.equal(p0, p0b) and torch.equal(v0, v0b), (n, count)
            for i in range(0, count, len(base)):
                m = min(len(base), count - i)
                for pp, vv, tol in ((p0, v0, 1e-3), (p2, v2, 3e-4)):
                    assert np.abs(pp[i:i + m].cpu().numpy() - ref_p[:m]).max() < tol, (n, count, i, tol)
                    assert np.abs(vv[i:i + m].cpu().numpy() - ref_v[:m]).max() < tol, (n, count, i, tol)
            assert float((p2 - p1).abs().max()) < 2e-4 and float((v2 - v1).abs().max()) < 2e-4
            np.testing.assert_allclose(p0.sum(dim=1).cpu().numpy(), 1.0, atol=1e-5)
        # device-side count
        count = 200
        pos = [base[i % len(base)] for i in range(count)]
        d_boards = torch.from_numpy(game.boards_from_states([p[0] for p in pos]).view(np.int64)).cuda()
        d_who = torch.tensor([p[1] for p in pos], dtype=torch.uint8, device="cuda")
        full_p, full_v = dn.forward_boards(d_boards, d_who, count, 0)
        probs = torch.full((count, n * n), -7.0, dtype=torch.float32, device="cuda")
        values = torch.full((count,), -7.0, dtype=torch.float32, device="cuda")
        d_count = torch.tensor([131], dtype=torch.int32, device="cuda")
        _cabi.check(_cabi.lib().caro_net_forward(dn.handle, game.game_kind, game.n, game.k, d_boards.data_ptr(), d_who.data_ptr(),
                                                 d_count.data_ptr(), count, probs.data_ptr(), values.data_ptr(), 0,
                                                 torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(probs[:131], full_p[:131]) and torch.equal(values[:131], full_v[:131])
        assert bool((probs[131:] == -7.0).all()) and bool((values[131:] == -7.0).all())
        dn.close()


def test_deeper_towers(torch_cuda):
    """BASELINE.json configs[3] names a "deep residual net" the reference does not define (lib/model.py has exactly five
    64-filter blocks).  `Net(..., blocks=N)` stacks N of the same blocks and every tower kernel reads the depth off the weight
    blob: 10 and 20 blocks on Connect4 (row-tiled bf16, split-precision and SIMT towers), 8 blocks on 15 x 15 (tap-per-MMA tower
    + tensor-core heads; biases beyond the sixth layer come from global memory) against PyTorch fp32, and a self-play step."""
    torch = torch_cuda
    from test_gpu_parity import _reference_outputs
    from harness import random_position
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import DeviceNet, Net
    rng = np.random.default_rng(8)

    def deep_net(game, blocks, seed):
        torch.manual_seed(seed)
        net = Net(game.obs_shape, game.action_space, blocks=blocks)
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.uniform_(-0.2, 0.2)
                    m.running_var.uniform_(0.7, 1.3)
                    m.weight.uniform_(0.6, 1.0)   # keeps the residual stream of a 20-block tower O(10)
                    m.bias.uniform_(-0.1, 0.1)
        return net.eval()

    for game, blocks, plies, count in ((ConnectFour(), 10, 30, 200), (ConnectFour(), 20, 30, 200), (TicTacToe(15, 5), 8, 40, 40)):
        og = oracle_for(game)
        net = deep_net(game, blocks, blocks)
        assert len([k for k in net.state_dict() if k.endswith(".0.weight") and k.startswith("conv_") and k[5].isdigit()]) == blocks
        pos = [random_position(og, rng, int(rng.integers(0, plies))) for _ in range(count)]
        states, players = [p[0] for p in pos], [p[1] for p in pos]
        ref_p, ref_v = _reference_outputs(game, net, states, players)
        dn = DeviceNet(net, game, precision="bf16")
        for impl, tol in ((1, 2e-4), (2, 5e-4), (0, 2e-3 if blocks > 10 else 1e-3)):
            p, v = dn.forward_states(states, players, impl=impl)
            dp, dv = np.abs(p.cpu().numpy() - ref_p).max(), np.abs(v.cpu().numpy() - ref_v).max()
            assert dp < tol and dv < tol, (type(game).__name__, blocks, impl, dp, dv)
        if blocks == 10:
            eng = SelfPlayEngine(game, 64, max_batch=8, node_capacity=1024, seed=4)
            eng.play(dn, dn, moves=3, count=8, batch=8, tau_plies=10, auto_restart=True)
            c = eng.counters()
            assert c["errors"] == 0 and c["plies"] == 64 * 3 and c["leaf_evals"] > 0
            eng.close()
        dn.close()


# --------------------------------------------------------------------------- replay ring -> SGD batch
def test_replay_gather_and_train_step_match_the_reference(torch_cuda, golden_train):
    """train.py:82-111 with the batch assembled on the device: the fixture's replay buffer is loaded into the engine's
    ring, `random.sample` (same seed as the reference run) picks the same rows, the CUDA gather kernel produces planes /
    pi / z equal to the oracle's encoding of those rows, the first round's losses match the reference to 1e-5 and the ten
    rounds' mean losses to 2e-3 (GPU fp32 convolutions, TF32 off, vs the reference's CPU run)."""
    torch = torch_cuda
    import torch.optim as optim
    from caro_ai_b200 import config as cfg, train as T
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import Net
    from helpers import oracle_game
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")

    class Rec:
        def __init__(self):
            self.rows = {}

        def track(self, name, value, step):
            self.rows[name] = float(value)

    for case in golden_train:
        game = ConnectFour() if case["game"] == "connect4" else TicTacToe(3, 3)
        og = oracle_game(case["game"])
        entries = [(s, p, pi, z) for s, p, pi, z in case["replay"]]
        eng = SelfPlayEngine(game, 4, max_batch=8, node_capacity=64, replay_capacity=1024, seed=1)
        eng.replay_load(entries[:250])
        eng.replay_load(entries[250:])
        assert eng.replay_live() == len(entries) == 600
        # (1) the gathered rows are the rows the reference sampled
        random.seed(case["sample_seed"])
        want = random.sample(collections.deque(entries), cfg.BATCH_SIZE)
        random.seed(case["sample_seed"])
        planes, pi, z = eng.replay_sample(cfg.BATCH_SIZE, random)
        np.testing.assert_array_equal(planes.cpu().numpy(), og.states_to_training_batch([b[0] for b in want], [b[1] for b in want]))
        np.testing.assert_array_equal(pi.cpu().numpy(), np.array([b[2] for b in want], dtype=np.float32))
        np.testing.assert_array_equal(z.cpu().numpy(), np.array([b[3] for b in want], dtype=np.float32))
        # (2) losses of the first round and of the whole call
        torch.manual_seed(case["net_seed"])
        net = Net(og.obs_shape, og.action_space).to(dev)
        net.train()
        with torch.no_grad():
            _, lv, lp = T.sgd_losses(net, planes, pi, z)
        # that forward updated BatchNorm's running statistics once: start again from the seed for the real call
        assert abs(lv.item() - case["round0"]["loss_value"]) < 1e-5 and abs(lp.item() - case["round0"]["loss_policy"]) < 1e-5
        torch.manual_seed(case["net_seed"])
        net = Net(og.obs_shape, og.action_space).to(dev)
        opt = optim.SGD(net.parameters(), lr=cfg.LEARNING_RATE, momentum=0.9)
        rec = Rec()
        random.seed(case["sample_seed"])
        means = T.train_neural_net(game, net, eng, opt, rec, 1, dev)
        for k, m in zip(("loss_total", "loss_value", "loss_policy"), means):
            assert abs(m - case["mean_losses"][k]) < 2e-3, (case["game"], k, m, case["mean_losses"][k])
            assert rec.rows[k] == m
        # ten SGD steps at lr 0.1 amplify the CPU / GPU rounding differences of the convolutions: the weights after the call
        # are compared through their summed magnitudes, 5 % (the per-round losses above are the tight check)
        sd = net.state_dict()
        for k, ref_sum in case["final_abs_sum"].items():
            got = float(sd[k].double().abs().sum())
            assert abs(got - ref_sum) <= 5e-2 * max(1.0, abs(ref_sum)), (k, got, ref_sum)
        # the reference-style deque goes through the same function (host states -> CUDA plane encoder)
        random.seed(case["sample_seed"])
        p2, pi2, z2 = T.sample_batch(game, collections.deque(entries), cfg.BATCH_SIZE, dev)
        assert torch.equal(p2, planes) and torch.equal(pi2, pi) and torch.equal(z2, z)
        eng.close()


def test_replay_symmetry_augmentation(torch_cuda, golden_train):
    """`replay_sample(augment=True)` (extension; the reference has no augmentation): every sampled row is written through a
    random board symmetry -- Connect4: column mirror; 3 x 3: the eight dihedral symmetries -- planes AND policy target
    transformed together, z untouched; checked against the same transforms applied on the host to the un-augmented rows."""
    torch = torch_cuda
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour, TicTacToe
    for case in golden_train:
        game = ConnectFour() if case["game"] == "connect4" else TicTacToe(3, 3)
        _, H, W = game.obs_shape
        A = game.action_space
        eng = SelfPlayEngine(game, 4, max_batch=8, node_capacity=64, replay_capacity=1024, seed=1)
        eng.replay_load([(s, p, pi, z) for s, p, pi, z in case["replay"]])
        random.seed(5)
        p0, pi0, z0 = (t.cpu().numpy() for t in eng.replay_sample(200, random))
        random.seed(5)
        p1, pi1, z1 = (t.cpu().numpy() for t in eng.replay_sample(200, random, augment=True))
        sym = eng._last_symmetry.cpu().numpy()
        assert set(sym.tolist()) == set(range(2 if A == W else 8))
        np.testing.assert_array_equal(z0, z1)

        def src(t, r, c):
            if t & 4:
                r, c = c, r
            return (H - 1 - r if t & 2 else r), (W - 1 - c if t & 1 else c)
        for i in range(200):
            t = int(sym[i])
            want = np.zeros_like(p0[i])
            for r in range(H):
                for c in range(W):
                    sr, sc = src(t, r, c)
                    want[:, r, c] = p0[i][:, sr, sc]
            np.testing.assert_array_equal(p1[i], want)
            if A == W:
                want_pi = pi0[i][::-1] if t & 1 else pi0[i]
            else:
                want_pi = np.array([pi0[i][src(t, a // W, a % W)[0] * W + src(t, a // W, a % W)[1]] for a in range(A)], dtype=np.float32)
            np.testing.assert_array_equal(pi1[i], want_pi)
        eng.close()


def test_selfplay_worker_keeps_one_engine_and_a_ring_of_several_steps(torch_cuda):
    """The trainer's self-play side: the same engine (same workspace pointer) serves every step, every step plays
    `games` fresh games to the end, the ring keeps the positions of the last `replay_steps` steps (nothing of the current
    step is evicted by the step itself), and a sampled batch has the network's shapes."""
    torch = torch_cuda
    from caro_ai_b200.game import TicTacToe
    from caro_ai_b200.model import DeviceNet, Net
    from caro_ai_b200.utils import SelfPlayWorker
    game = TicTacToe(3, 3)
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    w = SelfPlayWorker(game, 64, 6, 8, 3, replay_steps=2, min_replay=100, seed=7)
    assert w.replay_capacity == 2 * 64 * 9
    ptr = w.engine.workspace.data_ptr()
    seen = []
    total = 0
    for step in range(4):
        s = w.play_step(dn)
        assert s["games"] == 64 == s["wins"] + s["losses"] + s["draws"] and 64 * 5 <= s["plies"] <= 64 * 9
        total += s["plies"]
        assert w.engine.workspace.data_ptr() == ptr
        assert w.replay_len() == min(total, w.replay_capacity)
        seen.append(w.engine.roots()[0][:8])
    planes, pi, z = w.engine.replay_sample(256)
    assert tuple(planes.shape) == (256, 2, 3, 3) and tuple(pi.shape) == (256, 9) and tuple(z.shape) == (256,)
    assert bool(((z == 0) | (z == 1) | (z == -1)).all()) and torch.allclose(pi.sum(dim=1), torch.ones(256, device="cuda"), atol=1e-5)
    w.close()
    # the same worker with tree compaction (small arenas): the same games, position for position, in the replay ring
    from caro_ai_b200.game import ConnectFour
    c4 = ConnectFour()
    torch.manual_seed(0)
    d4 = DeviceNet(Net(c4.obs_shape, c4.action_space).eval(), c4)
    a = SelfPlayWorker(c4, 96, 8, 8, 10, replay_steps=1, min_replay=100, seed=5)
    b = SelfPlayWorker(c4, 96, 8, 8, 10, replay_steps=1, min_replay=100, seed=5, compact_tree=True)
    assert b.engine.cfg.node_capacity < a.engine.cfg.node_capacity
    for step in range(2):
        sa, sb = a.play_step(d4), b.play_step(d4)
        assert sa == sb and sa["games"] == 96
    assert a.replay_len() == b.replay_len()

    def rows(w):  # ring entries as a sorted list (games that finish in the same ply reserve their slots in no fixed order)
        n = w.replay_len()
        e = w.engine
        bd = e.region("replay_board")[:n].cpu().numpy().view(np.uint64).reshape(n, -1)
        pl = e.region("replay_player")[:n].cpu().numpy()
        pi = e.region("replay_pi")[:n].cpu().numpy()
        z = e.region("replay_z")[:n].cpu().numpy()
        return sorted((tuple(bd[i].tolist()), int(pl[i]), tuple(pi[i].tolist()), float(z[i])) for i in range(n))

    assert rows(a) == rows(b)
    a.close()
    b.close()
    d4.close()
    dn.close()


def test_train_main_trains_evaluates_and_checkpoints(torch_cuda, tmp_path, capsys):
    """`python -m caro_ai_b200.train -g 1` end to end on one GPU (train.py:165-217): two steps of 512 TicTacToe games fill the
    device ring past MIN_REPLAY_TO_TRAIN, so both steps run the ten SGD rounds, the arena is played after the second step, the
    reference's progress line is printed, and a promotion (if the arena grants one) writes `saves/<name>/best_001_00002.dat`
    in the reference's format."""
    torch = torch_cuda
    from caro_ai_b200 import train as train_cli
    from caro_ai_b200.game import TicTacToe
    from caro_ai_b200.model import load_checkpoint
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        random.seed(0)
        torch.manual_seed(0)
        assert train_cli.main(["-g", "1", "-n", "r2test", "--games", "512", "--max-steps", "2", "--evaluate-every", "2", "--augment"]) == 0
    finally:
        os.chdir(cwd)
    out = capsys.readouterr().out
    assert "Step 2, steps" in out and "replay " in out and "Net evaluated, win ratio = " in out
    ratio = float(out.split("Net evaluated, win ratio = ")[1].split()[0])
    saved = os.path.join(tmp_path, "saves", "r2test", "best_001_00002.dat")
    assert os.path.isdir(os.path.join(tmp_path, "saves", "r2test"))
    assert os.path.exists(saved) == (ratio > 0.60)
    if os.path.exists(saved):
        load_checkpoint(saved, TicTacToe(3, 3))


# --------------------------------------------------------------------------- cached ply graph life cycle
def test_ply_graph_is_not_replayed_for_other_engines_or_new_weights(torch_cuda):
    """`caro_engine_play_multi` replays one captured CUDA graph per ply.  The graph bakes in workspace pointers,
    dimensions and the tower's by-value constants, so it must be rebuilt when (a) the engines were destroyed and new ones
    (possibly at the same heap addresses, with another number of games) take their place, (b) the network's weights were
    updated in place.  Each pipelined run is compared bit for bit with single-engine `play()` runs of the same seeds."""
    torch = torch_cuda
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.model import DeviceNet, Net
    game = ConnectFour()
    torch.manual_seed(0)
    net_a = Net(game.obs_shape, game.action_space).eval()
    torch.manual_seed(1)
    net_b = Net(game.obs_shape, game.action_space).eval()
    dn = DeviceNet(net_a, game, precision="bf16")
    dn_ref = DeviceNet(net_a, game, precision="bf16")

    def run(sizes, seed):
        pair = [SelfPlayEngine(game, g, max_batch=8, node_capacity=1024, seed=seed + h) for h, g in enumerate(sizes)]
        SelfPlayEngine.play_multi(pair, dn, moves=3, count=6, batch=8, tau_plies=10, auto_restart=True)
        solo = [SelfPlayEngine(game, g, max_batch=8, node_capacity=1024, seed=seed + h) for h, g in enumerate(sizes)]
        for e in solo:
            e.play(dn_ref, dn_ref, moves=3, count=6, batch=8, tau_plies=10, auto_restart=True)
        torch.cuda.synchronize()
        for a, b in zip(pair, solo):
            assert a.counters() == b.counters() and a.counters()["errors"] == 0
            assert a.roots() == b.roots() and torch.equal(a.region("nodes"), b.region("nodes"))
        for e in pair + solo:
            e.close()

    run((96, 96), 40)
    run((96, 96), 40)     # same shapes, NEW engines: the old graph points into freed workspaces
    run((64, 160), 41)    # other sizes
    dn.update(net_b)      # weights rewritten in place: the captured constants are stale
    dn_ref.update(net_b)
    run((64, 160), 41)
    dn.close()
    dn_ref.close()


def test_facade_draws_fresh_noise_on_every_search_batch(torch_cuda):
    """lib/mcts.py:131-132 draws fresh Dirichlet noise for every descent.  The facade searches the same engine slot again
    and again (same game id, same ply, same side), so successive `search_batch` calls must address DIFFERENT Philox
    vectors: `caro_engine_search` takes the index of its first minibatch, two searches of 6 minibatches numbered 0..5 and
    6..11 grow exactly the tree of one search of 12, a search that restarts at 0 does not, and the facade passes its
    running minibatch count.  A module whose weights change between searches is re-folded (live weights, like the
    reference)."""
    torch = torch_cuda
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.mcts import MCTS
    from caro_ai_b200.model import DeviceNet, Net
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")
    engs = [SelfPlayEngine(game, 32, max_batch=8, node_capacity=256, seed=9) for _ in range(3)]
    engs[0].search(dn, 12, 8)
    engs[1].search(dn, 6, 8, first_minibatch=0)
    engs[1].search(dn, 6, 8, first_minibatch=6)
    engs[2].search(dn, 6, 8)
    engs[2].search(dn, 6, 8)  # numbered 0..5 again: the first search's noise vectors are replayed
    torch.cuda.synchronize()
    assert torch.equal(engs[0].region("nodes"), engs[1].region("nodes"))
    assert not torch.equal(engs[0].pool("N"), engs[2].pool("N"))
    for e in engs:
        e.close()
    t = MCTS(game, node_capacity=4096, seed=3)
    seen = []
    real = t._eng().search
    t._eng().search = lambda net, count, batch, impl=None, first_minibatch=0: (seen.append(first_minibatch),
                                                                             real(net, count, batch, impl, first_minibatch))[1]
    s0 = game.initial_state
    t.search_batch(12, 8, s0, 0, dn)
    t.search_batch(5, 8, s0, 0, dn)
    t.search_batch(3, 8, s0, 0, dn)
    assert seen == [0, 12, 17]
    net = Net(game.obs_shape, game.action_space).eval()
    c = MCTS(game, node_capacity=4096, seed=3)
    c.search_batch(4, 8, s0, 0, net)
    p_before = list(c.probs[s0])
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.5)
    c.clear()
    c.search_batch(4, 8, s0, 0, net)
    assert list(c.probs[s0]) != p_before
    dn.close()


def test_facade_reports_a_full_arena(torch_cuda):
    torch = torch_cuda
    from caro_ai_b200 import _cabi
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.mcts import MCTS
    from caro_ai_b200.model import DeviceNet, Net
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    t = MCTS(game, node_capacity=16, seed=1)
    t.search_batch(10, 8, game.initial_state, 0, dn)
    with pytest.raises(_cabi.CaroError):
        t.get_policy_value(game.initial_state)
    dn.close()


# --------------------------------------------------------------------------- shipped checkpoints
def test_shipped_checkpoints_load_and_later_generation_wins_its_arena(torch_cuda):
    """All four shipped checkpoints (saves/trained_connect4/best_025|026, saves/trained_tictactoe/best_004|005) load through
    the reference's `.dat` format.  Generation 026 was promoted over 025 by the reference's arena (train.py:120-149: games at
    search_batch(20, 16), tau = 0, promotion above 0.60): replayed here -- 2 x 1,000 games, both colours, fresh trees per game
    and side, the split-precision tower the precision check selects for these networks -- 026 must win that arena again
    (measured 0.61; the unmodified reference on the CPU, 160 games at search_batch(20, 8): 0.575 with eval-mode BatchNorm,
    0.67 with its train-mode BatchNorm).  The ordering is NOT robust to the search budget -- at play.py's 40 x 8 the older
    net wins 0.61, at 80 x 8 it is even (tools/strength_diag.py) -- so only the arena's own setting is
    asserted."""
    torch = torch_cuda
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import DeviceNet, load_checkpoint
    from caro_ai_b200.utils import play_games_batched
    ck = os.path.join(GOLDEN, "checkpoints")
    for name, game in [("tictactoe_best_004_00800.dat", TicTacToe(3, 3)), ("tictactoe_best_005_00900.dat", TicTacToe(3, 3))]:
        net = load_checkpoint(os.path.join(ck, name), game).eval()
        dn = DeviceNet(net, game)
        p, v = dn.forward_states([game.initial_state], [0])
        assert abs(float(p.sum().item()) - 1.0) < 1e-5 and abs(float(v[0].item())) <= 1.0
        dn.close()
    game = ConnectFour()
    new = DeviceNet(load_checkpoint(os.path.join(ck, "connect4_best_026_12000.dat"), game).eval(), game)
    old = DeviceNet(load_checkpoint(os.path.join(ck, "connect4_best_025_10600.dat"), game).eval(), game)
    assert new.precision == old.precision == "bf16x3"
    rounds = 1000
    a = play_games_batched(game, rounds, new, old, 0, 20, 16, trees_per_game=2, seed=1)
    b = play_games_batched(game, rounds, old, new, 0, 20, 16, trees_per_game=2, seed=2)
    assert a["games"] == b["games"] == rounds
    new_wins, old_wins = a["wins"] + b["losses"], a["losses"] + b["wins"]
    assert new_wins > 1.15 * old_wins, (a, b)
    new.close()
    old.close()


# --------------------------------------------------------------------------- throughput-mode extensions (not in the reference)
def _match(torch, game, dn, mode_a, mode_b, games, a_first, seed):
    """Engine A (search settings mode_a) plays player 0, engine B (mode_b) player 1, `games` games in lock-step (every
    game has the same side to move, so each ply is searched by one engine and the positions are copied to the other;
    both engines keep their own trees across the game).  Returns (points of A, games, leaf evals of A, leaf evals of B)."""
    from caro_ai_b200.engine import SelfPlayEngine
    engs = []
    for i, m in enumerate((mode_a, mode_b)):
        engs.append(SelfPlayEngine(game, games, max_batch=8, node_capacity=8192, seed=seed + i, virtual_loss=m["virtual_loss"]))
    first = 0 if a_first else 1
    for e in engs:
        e.reset(first_player=first)
    modes = (mode_a, mode_b)
    side = first
    for ply in range(42):
        e, o, m = engs[side], engs[1 - side], modes[side]
        e.search(dn, m["count"], 8, first_minibatch=ply * 64)
        e.advance(0, None, auto_restart=False, want_actions=False)
        for name in ("root_board", "root_player", "status", "ply"):
            o.region(name).copy_(e.region(name))
        side = 1 - side
        if int((e.region("status") == 0).sum().item()) == 0:
            break
    ca, cb = engs[0].counters(), engs[1].counters()
    assert ca["errors"] == 0 and cb["errors"] == 0
    wins_a = ca["wins_p0"] + cb["wins_p0"]
    wins_b = ca["wins_p1"] + cb["wins_p1"]
    draws = ca["draws"] + cb["draws"]
    assert wins_a + wins_b + draws == games
    for x in engs:
        x.close()
    return wins_a + 0.5 * draws, games, ca["leaf_evals"], cb["leaf_evals"]


def test_virtual_loss_spreads_descents_and_keeps_strength(torch_cuda):
    """CARO_FLAG_VIRTUAL_LOSS (extension, default off).  (1) Without it 30-55 % of the descents of a Connect4 search reach the
    network (lib/mcts.py:273-278 drops the duplicates; 54 % over the first six plies at 50 x 8, ~30 % at 100 x 8 in the middle
    game); with it >= 90 %.  (2) At (roughly) equal leaf evaluations per move --
    10 x 8 descents with virtual loss vs 40 x 8 without -- the virtual-loss search is not weaker: over 2 x 1,024 games of the
    shipped Connect4 checkpoint (both colours, tau = 0) it scores >= 45 % (50 % = parity; the verdict's +-3 % band needs more
    games than a test should play)."""
    torch = torch_cuda
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour
    from caro_ai_b200.model import DeviceNet, Net, load_checkpoint
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")
    frac = {}
    for vl in (False, True):
        eng = SelfPlayEngine(game, 1024, max_batch=8, node_capacity=8192, seed=5, virtual_loss=vl)
        eng.play(dn, dn, moves=6, count=50, batch=8, tau_plies=10, auto_restart=True)  # six plies: no game can have ended
        c = eng.counters()
        assert c["errors"] == 0 and c["descents"] == 1024 * 8 * 50 * 6
        frac[vl] = c["leaf_evals"] / c["descents"]
        # the tree invariants hold in both modes
        nodes = eng.region("node_count").cpu().numpy()
        assert int(nodes.sum()) == c["leaf_evals"]
        eng.close()
    assert frac[False] < 0.7 and frac[True] >= 0.9 and frac[True] > 1.4 * frac[False], frac
    dn.close()
    # the eight-lanes-per-game kernel (default for A <= 8) makes exactly the choices of the one-thread-per-game kernel
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    for mode in ("1", "0"):
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "vl_digest.py")], capture_output=True, text=True, timeout=600,
                             env=dict(os.environ, CARO_VL_GROUP=mode), cwd=root)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0].replace("hash", ""))
    assert digests[0] == digests[1], digests
    trained = DeviceNet(load_checkpoint(os.path.join(GOLDEN, "checkpoints", "connect4_best_026_12000.dat"), game).eval(), game)
    vl_mode, ref_mode = {"virtual_loss": True, "count": 10}, {"virtual_loss": False, "count": 40}
    p1, g1, ev_vl1, ev_ref1 = _match(torch, game, trained, vl_mode, ref_mode, 1024, True, 100)
    p2, g2, ev_ref2, ev_vl2 = _match(torch, game, trained, ref_mode, vl_mode, 1024, True, 200)
    score_vl = (p1 + (g2 - p2)) / (g1 + g2)
    evals_vl, evals_ref = ev_vl1 + ev_vl2, ev_ref1 + ev_ref2
    assert 0.6 < evals_vl / evals_ref < 1.6, (evals_vl, evals_ref)
    print("virtual loss vs reference search: score %.3f at %.2fx the leaf evaluations" % (score_vl, evals_vl / evals_ref))
    assert score_vl >= 0.45, (score_vl, evals_vl, evals_ref)
    trained.close()


def test_masked_priors_and_fresh_tree_flags(torch_cuda):
    """CARO_FLAG_MASK_PRIORS: every node's priors vanish on illegal moves and sum to one (the reference keeps the raw
    softmax).  CARO_FLAG_FRESH_TREE: a game's arena is emptied after each of its moves, so a capacity of ONE move's searches
    carries a game of any length (here 15 x 15 with a 2,048-node arena over 12 plies of 200 descents: no arena-full bit);
    CARO_FLAG_RECYCLE_TREE keeps the tree until it fills half the arena (reuse on most moves, still no overflow); with neither
    the persistent tree overflows this arena."""
    torch = torch_cuda
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import DeviceNet, Net
    game = ConnectFour()
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")
    eng = SelfPlayEngine(game, 64, max_batch=8, node_capacity=4096, seed=2, mask_priors=True)
    eng.play(dn, dn, moves=30, count=10, batch=8, tau_plies=10, auto_restart=False)  # deep into the games: full columns exist
    assert eng.counters()["errors"] == 0
    checked = 0
    for g in range(0, 64, 7):
        for s, n in eng.export_tree(g).items():
            legal = game.possible_moves(s)
            p = np.asarray(n["P"], dtype=np.float64)
            assert all(p[a] == 0 for a in range(7) if a not in legal) and abs(p.sum() - 1.0) < 1e-5
            checked += len(legal) < 7
    assert checked > 10
    eng.close()
    dn.close()
    caro = TicTacToe(15, 5)
    torch.manual_seed(0)
    dc = DeviceNet(Net(caro.obs_shape, caro.action_space).eval(), caro, precision="bf16")
    for mode in ("fresh", "recycle", "keep"):
        eng = SelfPlayEngine(caro, 32, max_batch=8, node_capacity=2048, seed=3, fresh_tree=mode == "fresh", recycle_tree=mode == "recycle")
        reused = 0
        for ply in range(12):
            eng.play(dc, dc, moves=1, count=25, batch=8, tau_plies=10, auto_restart=True)
            reused += int((eng.region("node_count") > 0).sum().item())
        c = eng.counters()
        assert bool(c["errors"] & 1) == (mode == "keep"), (mode, c)
        if mode == "fresh":
            assert reused == 0  # emptied by every advance
        if mode == "recycle":  # kept most of the time, never more than half an arena + one move's searches
            assert reused > 32 * 6 and int(eng.region("node_count").max().item()) <= 1024
        eng.close()
    dc.close()


def test_tree_compaction_changes_nothing_that_can_be_reached(torch_cuda):
    """CARO_FLAG_COMPACT_TREE drops, after every move, the nodes whose position can no longer occur and packs the arena.
    Against a twin engine that keeps everything (the reference's tree), with the same seeds and the real tower: the same
    moves, policies, roots and counters ply after ply (so every later search saw the same statistics), every surviving
    node bit-identical to its twin (N / W / Q / P / promotion flags), every twin node that contains the root position still
    present, far fewer nodes -- for Connect4 (one and two trees per game) and a 9 x 9 five-in-a-row board; and a 15 x 15
    game of 14 moves at 200 descents per move fits a 2,048-node arena that the keep-everything tree overflows."""
    torch = torch_cuda
    from caro_ai_b200.engine import SelfPlayEngine
    from caro_ai_b200.game import ConnectFour, TicTacToe
    from caro_ai_b200.model import DeviceNet, Net
    for game, tpg, G, cap, count, plies in ((ConnectFour(), 1, 256, 4096, 20, 14), (ConnectFour(), 2, 128, 4096, 12, 10),
                                            (TicTacToe(9, 5), 1, 48, 4096, 20, 10)):
        og = oracle_for(game)
        torch.manual_seed(0)
        dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game, precision="bf16")
        keep = SelfPlayEngine(game, G, trees_per_game=tpg, max_batch=8, node_capacity=cap, seed=11)
        comp = SelfPlayEngine(game, G, trees_per_game=tpg, max_batch=8, node_capacity=cap, seed=11, compact_tree=True)
        for ply in range(plies):
            acts = []
            for e in (keep, comp):
                e.search(dn, count, 8, first_minibatch=ply * count)
                pi, q, n = e.root_policy(2, 10)
                acts.append((pi.clone(), q.clone(), n.clone(), e.advance(10, None, auto_restart=True).clone()))
            for a, b in zip(acts[0], acts[1]):
                assert torch.equal(a, b), (game.obs_shape, tpg, ply)
            assert keep.roots() == comp.roots() and keep.counters() == comp.counters()
        assert comp.counters()["errors"] == 0
        nk, nc = keep.region("node_count").cpu().numpy(), comp.region("node_count").cpu().numpy()
        assert (nc <= nk).all() and nc.sum() < 0.7 * nk.sum(), (nk.sum(), nc.sum())
        roots, _ = comp.roots()
        for tree in range(0, G * tpg, max(1, G * tpg // 9)):
            tk, tc = keep.export_tree(tree), comp.export_tree(tree)
            root = roots[tree // tpg]
            rb = game.boards_from_states([root]).view(np.uint64)[0]
            for s, node in tc.items():
                twin = tk[s]
                assert node["N"] == twin["N"] and node["f32"] == twin["f32"]
                for k in ("W", "Q", "P"):
                    assert np.array_equal(node[k], twin[k]), (tree, s, k)
            for s in tk:  # everything that still contains the root position survived
                qb = game.boards_from_states([s]).view(np.uint64)[0]
                if isinstance(game, ConnectFour):
                    inside = (rb[0] & ~qb[0]) == 0 and (qb[1] & rb[0]) == rb[1]
                else:
                    inside = all((rb[i] & ~qb[i]) == 0 for i in range(8))
                assert (s in tc) == bool(inside), (tree, s)
        keep.close()
        comp.close()
        dn.close()
    caro = TicTacToe(15, 5)
    torch.manual_seed(0)
    dc = DeviceNet(Net(caro.obs_shape, caro.action_space).eval(), caro, precision="bf16")
    for compact in (True, False):
        eng = SelfPlayEngine(caro, 32, max_batch=8, node_capacity=2048, seed=3, compact_tree=compact)
        eng.play(dc, dc, moves=14, count=25, batch=8, tau_plies=10, auto_restart=True)
        assert bool(eng.counters()["errors"] & 1) == (not compact), (compact, eng.counters())
        eng.close()
    dc.close()
