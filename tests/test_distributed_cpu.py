"""CPU test of the N>1 plumbing: world_size 2 on gloo (weights broadcast + round-robin shard + result all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crisperwhisper_b200 import distributed as D
    from crisperwhisper_b200 import weights as Wt
    cfg = Wt.make_config(d_model=128, n_heads=2, enc_layers=1, dec_layers=1, ffn_dim=256, vocab=300, n_mels=80, eos_id=256,
                         no_timestamps_id=290, alignment_heads=[[0, 1]])
    pw = Wt.synthetic_weights(cfg, "cpu", seed=3) if rank == 0 else None
    got = D.broadcast_weights(pw, cfg, "cpu")
    ref = Wt.synthetic_weights(cfg, "cpu", seed=3)
    ok_w = torch.equal(got.arena_bf16, ref.arena_bf16) and torch.equal(got.arena_f32, ref.arena_f32)
    n_items = 7
    mine = D.shard_round_robin(n_items, rank, world)
    local = [(np.arange(i + 3, dtype=np.int64) + 10 * i, np.arange(i + 3, dtype=np.float32) * 0.02) for i in mine]
    full = D.gather_results(local, n_items, "cpu")
    ok_g = all(np.array_equal(full[i][0], np.arange(i + 3) + 10 * i) and
               np.allclose(full[i][1], np.arange(i + 3, dtype=np.float32) * 0.02) for i in range(n_items))
    q.put((rank, ok_w, ok_g, mine))
    dist.destroy_process_group()


def test_world2_gloo_broadcast_shard_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert res[0][3] == [0, 2, 4, 6] and res[1][3] == [1, 3, 5]
    assert all(r[1] and r[2] for r in res), res
