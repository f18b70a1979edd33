"""GPU parity tests for a19 (IndelRefine banded DP) through the C ABI: every group of segments captured from the reference
must produce exactly the blocks of the pinned oracle; all three kernel classes are exercised."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po
import irgen

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    import lra_b200
    c = lra_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["ir_ccs", "ir_ont", "ir_clr"])
def test_ir_dp_golden_segments(ctx, name):
    recs = po.read_ir_capture(os.path.join(GOLD, name + ".bin"))
    rg = irgen.groups_of_records(recs)
    r0 = recs[0]
    gb, expect = irgen.pack_groups(recs, rg, (r0["match"], r0["mismatch"], r0["indel"]))
    q = ctx.seq_upload(gb["q_arena"][:-16]); t = ctx.seq_upload(gb["t_arena"][:-16])
    r = ctx.indel_dp_batch(q, t, gb)
    q.free(); t.free()
    assert r["cells"] == int(sum((g["qE"] - g["qS"] + 1).sum() for _, g in rg))
    names = [s["name"] for s in ctx.kernel_stats()]
    assert any(n.startswith("ir_dp") for n in names)
    for j, e in enumerate(expect):
        assert r["n_blocks"][j] == len(e), (name, j)
        o = int(r["block_off"][j])
        assert (r["blocks"][o:o + len(e)] == e).all(), (name, j)


def test_ir_dp_errors(ctx):
    from lra_b200 import capi
    recs = po.read_ir_capture(os.path.join(GOLD, "ir_ccs.bin"))[:1]
    rg = irgen.groups_of_records(recs)[:8]
    gb, expect = irgen.pack_groups(recs, rg, (recs[0]["match"], recs[0]["mismatch"], recs[0]["indel"]))
    q = ctx.seq_upload(gb["q_arena"][:-16]); t = ctx.seq_upload(gb["t_arena"][:-16])
    with pytest.raises(capi.LraB200Error) as e:
        ctx.indel_dp_batch(q, t, gb, block_cap=2)
    assert e.value.code == capi.EOVERFLOW
    bad = dict(gb); bad["band"] = gb["band"].copy(); bad["band"][1] = bad["band"][0] - 5   # non-monotone qS
    with pytest.raises(capi.LraB200Error) as e:
        ctx.indel_dp_batch(q, t, bad)
    assert e.value.code == capi.EINVAL
    q.free(); t.free()
