"""GPU measurements of the BASELINE.json configurations other than the headline one (which is bench.py):

  1  TicTacToe 3,3,3 self-play, ~30 sims/move (train.py -g 1), beside the CPU oracle port on the host
  3  Connect4 self-play + the SGD step of the training loop (train.py:82-111) on this GPU
  4  Caro 15,15,5 self-play, 1600 sims/move
  5  tournament: random-init Connect4 checkpoints through the .dat format, every ordered pair, tau = 0

One JSON line per configuration on stdout.  Usage: python tools/config_bench.py [1,3,4,5]"""
import collections
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch

from caro_ai_b200 import config as cfg
from caro_ai_b200.engine import SelfPlayEngine
from caro_ai_b200.game import ConnectFour, TicTacToe
from caro_ai_b200.model import DeviceNet, Net, load_checkpoint, save_checkpoint
from caro_ai_b200.utils import play_games_batched


def timed_plies(eng, dn, plies, count, batch, tau):
    """Lock-step self-play with re-seating: returns (seconds, counter deltas) over `plies` plies, CUDA-event timed."""
    eng.play(dn, dn, moves=2, count=count, batch=batch, tau_plies=tau, auto_restart=True)
    torch.cuda.synchronize()
    c0 = eng.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.play(dn, dn, moves=plies, count=count, batch=batch, tau_plies=tau, auto_restart=True)
    e1.record()
    torch.cuda.synchronize()
    c1 = eng.counters()
    assert c1["errors"] == 0
    return e0.elapsed_time(e1) / 1e3, {k: c1[k] - c0[k] for k in c1}


def config1():
    game = TicTacToe(3, 3)
    torch.manual_seed(0)
    net = Net(game.obs_shape, game.action_space).eval()
    dn = DeviceNet(net, game)
    out = {"config": 1, "workload": "TicTacToe 3,3,3 self-play, random-init 5x64 net, tau=1 for 10 plies, 4096 concurrent games"}
    for count, batch in ((4, 8), (30, 1)):
        eng = SelfPlayEngine(game, 4096, max_batch=batch, node_capacity=2048, seed=1)
        sec, d = timed_plies(eng, dn, 60, count, batch, cfg.STEPS_BEFORE_TAU_0)
        out["search_batch(%d,%d)" % (count, batch)] = {"games_per_sec": d["games"] / sec, "leaf_evals_per_sec": d["leaf_evals"] / sec,
                                                       "plies_per_sec": d["plies"] / sec}
        eng.close()
    # CPU oracle port of the reference on one host core: 20 games with search_batch(4, 8)
    from oracle.games import MNKOracle
    from oracle.mcts import OracleMCTS
    from oracle.net import OracleNet
    og = MNKOracle(3, 3)
    torch.set_num_threads(1)
    onet = OracleNet(og.obs_shape, og.action_space)
    rows = {"n": 0}
    fwd = onet.forward

    def counting(x):
        rows["n"] += int(x.shape[0])
        return fwd(x)

    onet.forward = counting
    np.random.seed(0)
    t0 = time.perf_counter()
    games = 0
    while games < 20:
        state, cur, tree, ply = og.initial_state, int(np.random.choice(2)), OracleMCTS(og), 0
        while True:
            tree.search_batch(4, 8, state, cur, onet)
            pi, _ = tree.get_policy_value(state, tau=1 if ply < cfg.STEPS_BEFORE_TAU_0 else 0)
            a = int(np.random.choice(og.action_space, p=pi))
            state, won = og.move(state, a, cur)
            cur, ply = 1 - cur, ply + 1
            if won or not og.possible_moves(state):
                break
        games += 1
    dt = time.perf_counter() - t0
    out["cpu_port_search_batch(4,8)"] = {"games_per_sec": games / dt, "leaf_evals_per_sec": rows["n"] / dt, "cores": 1,
                                         "sample": "20 games of the oracle port, 1 torch thread"}
    dn.close()
    return out


def config3():
    """One outer step of the training loop at train.py's own search setting (10 x 8) plus the 10 SGD rounds, on this rank's
    GPU (under torchrun every rank runs it and the collectives of the step are real; rank 0 prints)."""
    import torch.distributed as dist
    import torch.optim as optim
    from caro_ai_b200 import distributed as D, train as T
    from caro_ai_b200.utils import SelfPlayWorker
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    rank, ws = D.world()
    game = ConnectFour()
    torch.manual_seed(0)
    device = torch.device("cuda", torch.cuda.current_device())
    net = Net(game.obs_shape, game.action_space).to(device)
    D.broadcast_state_dict(net)
    best = DeviceNet(net, game)

    class _TB:
        def track(self, *a, **k):
            pass

    games = 4096
    worker = SelfPlayWorker(game, games, cfg.MCTS_SEARCHES, cfg.MCTS_BATCH_SIZE, cfg.STEPS_BEFORE_TAU_0, replay_steps=2,
                            min_replay=cfg.REPLAY_BUFFER, seed=1000003 * rank + 17)
    T.self_play(worker, best, _TB(), 0)  # warm-up step (workspace first touched)
    torch.cuda.synchronize()
    stats, sp = T.self_play(worker, best, _TB(), 1)
    opt = optim.SGD(net.parameters(), lr=cfg.LEARNING_RATE, momentum=0.9)
    bucket = D.FlatGradients(net.parameters())
    T.train_neural_net(game, net, worker.engine, opt, _TB(), 1, device, bucket)  # warm-up (cuDNN autotune)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    losses = T.train_neural_net(game, net, worker.engine, opt, _TB(), 2, device, bucket)
    torch.cuda.synchronize()
    sgd = time.perf_counter() - t1
    n_param = sum(p.numel() for p in net.parameters())
    tot = torch.tensor([games / sp, stats["leaf_evals"] / sp], dtype=torch.float64, device=device)
    if ws > 1:
        dist.all_reduce(tot)
    out = {"config": 3, "ranks": ws,
           "workload": "Connect4 training step: %d self-play games per rank at search_batch(%d,%d) played to the end on ONE persistent engine, "
                       "then %d SGD rounds of %d rows drawn from the device replay rings (all-gather of %d rows per rank, one flattened "
                       "NCCL all-reduce of %d fp32 gradients per round)" % (games, cfg.MCTS_SEARCHES, cfg.MCTS_BATCH_SIZE, cfg.TRAIN_ROUNDS,
                                                                         cfg.BATCH_SIZE, cfg.BATCH_SIZE // ws, n_param),
           "self_play_games_per_sec": float(tot[0]), "self_play_leaf_evals_per_sec": float(tot[1]), "self_play_precision": best.precision,
           "replay_positions_per_rank": worker.replay_len(), "replay_capacity_per_rank": worker.replay_capacity,
           "sgd_ms_per_round": 1e3 * sgd / cfg.TRAIN_ROUNDS, "loss_total": losses[0], "gradient_bytes": 4 * n_param}
    worker.close()
    best.close()
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None


def config4(games=128, parts=1):
    game = TicTacToe(15, 5)
    torch.manual_seed(0)
    dn = DeviceNet(Net(game.obs_shape, game.action_space).eval(), game)
    if parts == 1:
        eng = SelfPlayEngine(game, games, max_batch=8, node_capacity=16384, seed=1)
        sec, d = timed_plies(eng, dn, 4, 200, 8, cfg.STEPS_BEFORE_TAU_0)
        eng.close()
    else:  # the parts pipeline (caro_engine_play_multi): tree kernels of one part under the other parts' towers
        engs = [SelfPlayEngine(game, games // parts, max_batch=8, node_capacity=16384, seed=1 + h) for h in range(parts)]

        def run(n):
            SelfPlayEngine.play_multi(engs, dn, moves=n, count=200, batch=8, tau_plies=cfg.STEPS_BEFORE_TAU_0, auto_restart=True)

        def tot():
            t = {}
            for e in engs:
                for k, v in e.counters().items():
                    t[k] = t.get(k, 0) + v
            return t

        run(2)
        torch.cuda.synchronize()
        c0 = tot()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(4)
        e1.record()
        torch.cuda.synchronize()
        c1 = tot()
        assert c1["errors"] == 0
        sec, d = e0.elapsed_time(e1) / 1e3, {k: c1[k] - c0[k] for k in c1}
        for e in engs:
            e.close()
    dn.close()
    return {"config": 4, "workload": "Caro 15,15,5 self-play, search_batch(200,8) = 1600 sims/move, %d concurrent games, reference-shape 5x64 "
                                     "net (tap-per-MMA tcgen05 kernel: boards larger than 6x7), 4 plies after 2 warm-up plies, %d pipeline part(s)" % (games, parts),
            "leaf_evals_per_sec": d["leaf_evals"] / sec, "plies_per_sec": d["plies"] / sec, "descents_per_sec": d["descents"] / sec,
            "tflops_useful": d["leaf_evals"] * 83760340 / sec / 1e12}


def config5(n_ckpt=8, rounds=1000):
    """BASELINE configs[4] at its stated size: 8 checkpoints through the `.dat` format, 56 ordered pairs x 1,000 games,
    search_batch(40,8), tau = 0, fresh trees per game and side -- through the play.py CLI twin (caro_ai_b200.play.main)."""
    import contextlib
    import io
    from caro_ai_b200 import play as play_cli
    game = ConnectFour()
    tmp = tempfile.mkdtemp()
    paths = []
    for i in range(n_ckpt):
        torch.manual_seed(100 + i)
        p = os.path.join(tmp, "net_%d.dat" % i)
        save_checkpoint(Net(game.obs_shape, game.action_space), p)
        paths.append(p)
    torch.cuda.synchronize()
    out, err = io.StringIO(), io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        play_cli.main(["-g", "0", "-r", str(rounds)] + paths)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lines = [l for l in out.getvalue().splitlines() if " vs " in l]
    table, total = {}, 0
    for l in lines:
        names, res = l.split(" -> ")
        a, b = [os.path.basename(x).replace("net_", "").replace(".dat", "") for x in names.split(" vs ")]
        w, lo, d = [int(x.split("=")[1].rstrip(",")) for x in res.split()]
        table["%s-%s" % (a, b)] = [w, lo, d]
        total += w + lo + d
    assert len(lines) == n_ckpt * (n_ckpt - 1) and total == len(lines) * rounds
    speeds = [float(l.split()[1]) for l in err.getvalue().splitlines() if l.startswith("Speed")]
    return {"config": 5, "workload": "tournament: %d random-init Connect4 checkpoints saved/loaded as .dat, %d ordered pairs x %d games, "
                                     "search_batch(40,8), tau=0, two trees per game (play.py semantics), through caro_ai_b200.play.main; "
                                     "wall clock incl. checkpoint loading, precision calibration and engine set-up per pair"
                                     % (n_ckpt, len(lines), rounds),
            "games_per_sec": total / dt, "games": total, "seconds": dt, "pair_games_per_sec_median": sorted(speeds)[len(speeds) // 2],
            "leaderboard": out.getvalue().split("Leaderboard:")[1].strip().splitlines(), "w_l_d": table}


def main():
    which = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 3, 4, 5]
    fns = {1: config1, 3: config3, 4: config4, 5: config5}
    for k in which:
        if k == 4 and len(sys.argv) > 2:
            print(json.dumps(config4(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 1)), flush=True)
        else:
            res = fns[k]()
            if res is not None:
                print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
