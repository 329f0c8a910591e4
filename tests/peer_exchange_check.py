"""Run under torchrun by tests/test_gpu_system.py: PeerExchange.gather must reproduce the NCCL all-gather bit for bit,
over several steps (both blocks of the double buffer), and the sharded evaluation on the gathered arrays must add up."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coulombgalore_b200 as cg  # noqa: E402
from coulombgalore_b200.distributed import HaloExchange, PeerExchange, ShardedStep, owned_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_side = 12
    N = n_side ** 3
    ev = cg.BatchedEvaluator(cg.Wolf(12.0, 0.25), local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ev.set_stream(stream.cuda_stream)
    ev.set_shard(rank, world)
    arr = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(4)]
    ev.generate_lattice_device(n_side, 3.0, 99, arr[0], arr[1], arr[2], arr[3])
    lo, hi = owned_range(N, rank, world)
    peer = PeerExchange.create(ev, N, 4, dev)
    assert isinstance(peer, PeerExchange), getattr(peer, "fallback_reason", None)
    nccl = ShardedStep(N, 4, dev)
    L = n_side * 3.0
    f = torch.zeros((N, 3), dtype=torch.float64, device=dev)
    sc = torch.zeros(2, dtype=torch.float64, device=dev)
    for step in range(5):
        owned = [((a + 0.01 * step) % L if k < 3 else a)[lo:hi].clone() for k, a in enumerate(arr)]  # positions move every step
        got = peer.gather(owned)
        ref = nccl.gather(owned)
        stream.synchronize()
        for g, r in zip(got, ref):
            assert torch.equal(g, r), "step %d" % step
        ev.set_system_device(got[0], got[1], got[2], got[3], None, [L] * 3, True)
        f.zero_()
        ev.eval_device(cg.MODE_ION, cg.ENERGY | cg.FORCE, force=f, scalars=sc)
        peer.reduce_scalars(sc)
        dist.all_reduce(f)
        stream.synchronize()
        if step == 0:
            single = cg.BatchedEvaluator(cg.Wolf(12.0, 0.25), local)
            single.set_system(*[t.cpu().numpy() for t in ref], None, [L] * 3, True)
            want = single.eval(cg.MODE_ION, cg.ENERGY | cg.FORCE)
            assert int(sc[1].item()) == want["pairs"] and abs(sc[0].item() - want["energy"]) <= 1e-12 * abs(want["energy"]) + 1e-9
            assert np.max(np.abs(f.cpu().numpy() - want["force"])) <= 1e-12 * np.max(np.abs(want["force"]))
            single.close()
    peer.close()
    ev.close()

    # ---- halo exchange: every rank receives only the particles of its window; results must not change ----
    n_side, L = 32, 96.0  # slab 48 + 2 x (cut-off + slack layers) < 96: a 2-rank window is a true subset
    N = n_side ** 3
    pot = cg.ReactionField(12.0, 80.0, 1.0, False)
    ev = cg.BatchedEvaluator(pot, local)
    ev.set_stream(stream.cuda_stream)
    ev.set_shard(rank, world)
    arr = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(7)]
    ev.generate_lattice_device(n_side, 3.0, 7, arr[0], arr[1], arr[2], arr[3], (arr[4], arr[5], arr[6]))
    lo, hi = owned_range(N, rank, world)
    halo = HaloExchange(ev, N, [L] * 3, 7, dev)
    single = cg.BatchedEvaluator(pot, local)
    single.set_system(*[t.cpu().numpy() for t in arr[:4]], tuple(t.cpu().numpy() for t in arr[4:]), [L] * 3, True)
    w = cg.ENERGY | cg.FORCE | cg.FIELD
    want = single.eval(cg.MODE_MULTIPOLE, w)
    single.close()
    received = []
    for step in range(6):  # both blocks several times: the publish / acknowledge words of the device-side ordering wrap around
        owned = torch.stack([a[lo:hi] for a in arr])
        sub, ids, n_out = halo.gather(owned)
        halo.set_system(ev, sub, has_q=True, has_mu=True)
        f = torch.zeros((sub[0].numel(), 3), dtype=torch.float64, device=dev)
        e = torch.zeros_like(f)
        sc2 = torch.zeros(2, dtype=torch.float64, device=dev)
        ev.eval_device(cg.MODE_MULTIPOLE, w, force=f, field=e, scalars=sc2)
        stream.synchronize()
        n_got = halo.check_overflow()
        received.append(n_got)
        full_f = torch.zeros((N, 3), dtype=torch.float64, device=dev)
        full_e = torch.zeros_like(full_f)
        idx = ids[:n_got].long()
        assert idx.unique().numel() == n_got  # no particle twice
        full_f[idx] = f[:n_got]
        full_e[idx] = e[:n_got]
        # cg_halo_collect_owned: the rows of this rank's own particles that it received, in the order of the arrays it published;
        # owned particles outside its window (ownership is by index here, not by position) are not touched
        own_f = torch.full((hi - lo, 3), float("nan"), dtype=torch.float64, device=dev)
        own_e = torch.full_like(own_f, float("nan"))
        halo.collect_owned([f, e], [own_f, own_e])
        stream.synchronize()
        got_mask = torch.zeros(N, dtype=torch.bool, device=dev)
        got_mask[idx] = True
        mine = got_mask[lo:hi]
        assert torch.equal(own_f[mine], full_f[lo:hi][mine]) and torch.equal(own_e[mine], full_e[lo:hi][mine]), "collect_owned, step %d" % step
        assert torch.isnan(own_f[~mine]).all()
        # cg_collect_evaluated: the particles this rank evaluated (its block of cell rows), compacted with their global indices;
        # all ranks together return every particle exactly once
        rows = ev.collect_rows_bound()
        ce_f = torch.full((rows, 3), float("nan"), dtype=torch.float64, device=dev)
        ce_e = torch.full_like(ce_f, float("nan"))
        ce_id = torch.full((rows,), -2.0, dtype=torch.float64, device=dev)
        ev.collect_evaluated([f, e], [ce_f, ce_e], ce_id, ids=halo.ids())
        stream.synchronize()
        assert (ce_id != -2.0).all()
        valid = ce_id >= 0
        cid = ce_id[valid].long()
        assert cid.unique().numel() == cid.numel()
        assert torch.equal(ce_f[valid], full_f[cid]) and torch.equal(ce_e[valid], full_e[cid]), "collect_evaluated, step %d" % step
        seen = torch.zeros(N, dtype=torch.float64, device=dev)
        seen[cid] = 1.0
        assert (full_f[seen == 0] == 0).all()  # what was not collected was not evaluated here
        dist.all_reduce(seen)
        assert (seen == 1.0).all(), "every particle is evaluated by exactly one rank"
        dist.all_reduce(full_f)
        dist.all_reduce(full_e)
        dist.all_reduce(sc2)
        assert int(sc2[1].item()) == want["pairs"], (int(sc2[1].item()), want["pairs"])
        assert abs(sc2[0].item() - want["energy"]) <= 1e-12 * abs(want["energy"]) + 1e-9
        assert np.max(np.abs(full_f.cpu().numpy() - want["force"])) <= 1e-12 * np.max(np.abs(want["force"]))
        assert np.max(np.abs(full_e.cpu().numpy() - want["field"])) <= 1e-12 * np.max(np.abs(want["field"]))
    if world > 1:
        assert received[0] < N, "a rank of a multi-rank run must receive a true subset"
    halo.close()
    ev.close()
    dist.barrier()
    if rank == 0:
        print("PEER_EXCHANGE_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
