"""N>1 host logic on CPU: two gloo ranks each hold one base shard, all-gather their 640-byte
partials (the exchange bench.py does over NCCL) and rank 0 assembles with fb_prove_finish."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import fawkes_crypto_b200 as fb
    from tests.dist_helpers import shard_partials
    from tests.util import fr_np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "synth_rows40.json")))
    mine = torch.frombuffer(bytearray(shard_partials(golden, rank, world)), dtype=torch.uint8)
    gathered = [torch.empty(640, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok = None
    if rank == 0:
        pb = bytes.fromhex(golden["bellman_params_hex"])
        parts = np.concatenate([g.numpy() for g in gathered])
        r, s = fr_np([int(golden["r"], 16)])[0], fr_np([int(golden["s"], 16)])[0]
        out = np.zeros(256, dtype=np.uint8)
        fb.native.check(fb.native.lib.fb_prove_finish(fb.native.ptr(pb), len(pb), parts.ctypes.data, world,
                                                      r.ctypes.data, s.ctypes.data, out.ctypes.data))
        ok = out.tobytes().hex() == golden["proof_raw_hex"]
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_two_rank_partial_exchange_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0] is True
