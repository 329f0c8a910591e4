"""Parity at the BASELINE sizes (full N): the forces / fields of an i-slice of >= 4096 particles
against the CPU oracle, through the same code path bench.py times (bench.Workload + bench.parity_slice), plus the
size-independent properties the domain offers at that size: the ordered in-cutoff pair count equals the count obtained by
summing over the shards of a two-way split, and the total force of a periodic system vanishes.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["1", "2", "3", "4", "5"])
def test_baseline_config_at_full_size(name, oracle_kind):
    import torch
    import bench
    cfg = bench.configs()[name]
    dev = torch.device("cuda", 0)
    # one dedicated stream for all evaluators, as in bench.py (cg_set_stream(NULL) would give every handle its own stream, and
    # the schemes of config 2 write the same output buffer one after the other)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    w = bench.Workload(name, cfg, 0, 1, 0, dev, stream)
    try:
        assert w.N == {"1": 4096, "2": 1000000, "3": 262144, "4": 4096000, "5": 16777216}[name]
        p = bench.parity_slice(w, oracle_kind)
        assert p["ok"] and p["worst"] <= 1.0, p
        for res in p["schemes"].values():
            assert res["particles"] >= 4096
        torch.cuda.synchronize()
        pairs = {k: w.scal[k][1].item() for k in w.names}
        assert all(v > 0 and v == int(v) for v in pairs.values())
        # Newton's third law over the whole periodic system: the forces sum to zero (to the rounding of ~N * neighbours terms)
        if cfg["pbc"] and w.force is not None:
            import coulombgalore_b200 as cg
            w.step()
            torch.cuda.synchronize()
            f = w.force
            if cfg["mode"] != cg.MODE_MULTIPOLE:  # the reference's multipole force mixes forces on different particles
                tot = f.sum(dim=0).abs().max().item()
                scale = f.abs().sum().item()
                assert tot <= 1e-12 * scale, (tot, scale)
        # the same evaluation as two shards: identical pair count; identical forces bit for bit (fixed summation order) with
        # whole groups as work items (CG_TAIL=0), to rounding with the tail split (the groups of the last round are summed
        # over row slices, and which groups those are depends on the shard)
        if name in ("2", "3"):
            import os
            import coulombgalore_b200 as cg
            ref_force = w.force.clone()
            if name == "3":
                os.environ["CG_TAIL"] = "0"
                try:
                    w.step()
                    torch.cuda.synchronize()
                    ref_whole = w.force.clone()
                    assert (ref_whole - ref_force).abs().max().item() <= 1e-12 * ref_force.abs().max().item()
                    out0 = torch.zeros_like(w.force)
                    w.lead.set_shard(0, 2)
                    w._set_system([w.full[k] for k in w.use])
                    w.lead.eval_device(cfg["mode"], cfg["what"], force=out0, field=w.field, scalars=torch.zeros(2, dtype=torch.float64, device=dev))
                    torch.cuda.synchronize()
                    written = out0.abs().sum(dim=1) > 0
                    assert written.any() and torch.equal(out0[written], ref_whole[written])
                finally:
                    del os.environ["CG_TAIL"]
                    w.lead.set_shard(0, 1)
            total = 0.0
            out = torch.zeros_like(w.force)
            for r in range(2):
                w.lead.set_shard(r, 2)
                out.zero_()
                sc = torch.zeros(2, dtype=torch.float64, device=dev)
                w._set_system([w.full[k] for k in w.use])
                w.lead.eval_device(cfg["mode"], cfg["what"], force=out, field=w.field, scalars=sc)
                torch.cuda.synchronize()
                total += sc[1].item()
                written = out.abs().sum(dim=1) > 0
                if name == "2":  # the buffers hold the LAST scheme of the config (Fennell); compare the lead scheme separately
                    continue
                assert (out[written] - ref_force[written]).abs().max().item() <= 1e-12 * ref_force.abs().max().item()
            w.lead.set_shard(0, 1)
            assert total == pairs[w.names[0]]
    finally:
        w.close()
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
