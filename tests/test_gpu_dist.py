"""Two real GPUs: MSM base shards + R1CS rows + four-step NTT over NCCL == the single-GPU proof.
Skipped on a one-GPU box (the layouts are covered there by test_distributed_h_pipeline_layouts and
test_sharded_prove_equals_unsharded)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["FB_DIST_NTT_MIN_LOG"] = "16"
    import ctypes as C
    import torch
    import torch.distributed as dist
    import fawkes_crypto_b200 as fb
    lib = fb.native.lib
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = fb.Context(rank)
    idbuf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        fb.native.check(lib.fb_dist_unique_id(idbuf.ctypes.data))
    idt = torch.from_numpy(idbuf).cuda()
    dist.broadcast(idt, src=0)
    idbuf = idt.cpu().numpy()
    fb.native.check(lib.fb_dist_init(ctx.handle, rank, world, idbuf.ctypes.data))
    seed = 0xFA3CE50000 + 1616
    circ = fb.Circuit.synthetic(1 << 16, seed)
    td = np.zeros((7, 4), dtype=np.uint64)
    fb.native.check(lib.fb_synth_trapdoor(seed, td.ctypes.data))
    tdi = [fb.groth16.fr_unraw(x) for x in td]
    params = fb.setup(circ, ctx, trapdoor=tdi[:5])
    wi, wa = circ.witness()
    # the key of this rank comes from the SHARDED setup (fb_setup_shard: only this rank's points are generated, the
    # rest of the byte string is points at infinity); the full setup above is kept for the single-GPU reference
    params_sh = fb.setup(circ, ctx, trapdoor=tdi[:5], shard=rank, nshards=world)
    assert len(params_sh.bellman_bytes) == len(params.bellman_bytes)
    assert bytes(params_sh.bellman_bytes[:580]) == bytes(params.bellman_bytes[:580])      # same verifying key
    pk = C.c_void_p()
    pb = params_sh.bellman_bytes
    fb.native.check(lib.fb_pk_load_shard(ctx.handle, fb.native.ptr(pb), len(pb), circ.handle, 1, rank, world, C.byref(pk)))
    # library-side collective prove: every rank calls fb_prove on its shard of the key; the 640-byte partial sums
    # travel over the library's own NCCL communicator and EVERY rank returns the proof
    r, s = fb.groth16.fr_raw(tdi[5]), fb.groth16.fr_raw(tdi[6])
    out = np.zeros(256, dtype=np.uint8)
    fb.native.check(lib.fb_prove(ctx.handle, pk, wi.ctypes.data, wi.shape[0], wa.ctypes.data, wa.shape[0],
                                 r.ctypes.data, s.ctypes.data, out.ctypes.data, None))
    # same through the device-resident entry point
    w_dev = torch.from_numpy(np.concatenate([wi, wa]).view(np.int64)).cuda()
    out_dev = np.zeros(256, dtype=np.uint8)
    fb.native.check(lib.fb_prove_device(ctx.handle, pk, w_dev.data_ptr(), r.ctypes.data, s.ctypes.data,
                                        out_dev.ctypes.data))
    # and with the caller's own transport: fb_prove_partial + all-gather + fb_prove_finish
    partial = np.zeros(640, dtype=np.uint8)
    fb.native.check(lib.fb_prove_partial(ctx.handle, pk, wi.ctypes.data, wi.shape[0], wa.ctypes.data, wa.shape[0],
                                         partial.ctypes.data))
    gathered = torch.empty((world, 640), dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(gathered, torch.from_numpy(partial).cuda())
    parts = gathered.cpu().numpy()
    out_manual = np.zeros(256, dtype=np.uint8)
    fb.native.check(lib.fb_prove_finish(fb.native.ptr(pb), 580, parts.ctypes.data, world, r.ctypes.data,
                                        s.ctypes.data, out_manual.ctypes.data))
    _, ref = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)   # unsharded key, same GPU
    ok = bool(out.tobytes() == ref.to_raw() and out_dev.tobytes() == ref.to_raw() and
              out_manual.tobytes() == ref.to_raw()) and fb.verify(params.get_vk(), ref, wi[1:])
    dist.barrier()
    lib.fb_pk_free(pk)
    dist.destroy_process_group()
    q.put((rank, ok))


def test_two_gpu_distributed_prove():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0] is True and res[1] is True   # every rank holds the proof
