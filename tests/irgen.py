"""Builds IndelRefine DP-group batches for the kernel tests from segments (captured from the reference) by running the
oracle restatement with its group dump (oracle/indel_refine.c: lra_oracle_indel_refine_groups)."""
import numpy as np
from oracle import pyoracle as po


def groups_of_records(recs):
    """-> list of (record index, group dict); the group's expected blocks come from the pinned restatement."""
    out = []
    for i, r in enumerate(recs):
        bo, st, groups = po.indel_refine_groups_port(r["read"], r["twin"], r["t_win_off"], r["contig_len"], r["blocks_in"], r["k"],
                                                     r["match"], r["mismatch"], r["indel"], r["end_align"])
        assert st == 0 and bo.shape == r["blocks_out"].shape and (bo == r["blocks_out"]).all()
        out += [(i, g) for g in groups]
    return out


def pack_groups(recs, rec_groups, scoring):
    """One batch (one scoring) over several segments: read strands and target windows are concatenated into two arenas.
    Block coordinates stay in the reference's frames: q relative to the read strand, t relative to the contig start, so
    t_base = (arena offset of the window) - t_win_off may wrap below zero; uint32 arithmetic makes t_base + tPos right."""
    used = sorted(set(i for i, _ in rec_groups))
    q_parts, t_parts, q_at, t_at = [], [], {}, {}
    qo = to = 0
    for i in used:
        q_at[i] = qo; q_parts.append(recs[i]["read"]); qo += len(recs[i]["read"])
        t_at[i] = to; t_parts.append(recs[i]["twin"]); to += len(recs[i]["twin"])
    q_arena = np.frombuffer(b"".join(q_parts) + b"N" * 16, dtype=np.uint8).copy()
    t_arena = np.frombuffer(b"".join(t_parts) + b"N" * 16, dtype=np.uint8).copy()
    n = len(rec_groups)
    gb = dict(q_arena=q_arena, t_arena=t_arena, q_base=np.zeros(n, np.uint32), t_base=np.zeros(n, np.uint32),
              q_start=np.zeros(n, np.int32), t_start=np.zeros(n, np.int32), t_len=np.zeros(n, np.int32),
              q_seq_len=np.zeros(n, np.int32), t_seq_len=np.zeros(n, np.int32), band_off=np.zeros(n, np.uint32),
              match=scoring[0], mismatch=scoring[1], indel=scoring[2])
    band, boff, expect = [], 0, []
    for j, (i, g) in enumerate(rec_groups):
        gb["q_base"][j] = q_at[i]
        gb["t_base"][j] = np.uint32((t_at[i] - recs[i]["t_win_off"]) & 0xFFFFFFFF)
        gb["q_start"][j] = g["qStart"]; gb["t_start"][j] = g["tStart"]; gb["t_len"][j] = g["tLen"]
        gb["q_seq_len"][j] = g["qSeqLen"]; gb["t_seq_len"][j] = g["tSeqLen"]; gb["band_off"][j] = boff
        band += [g["qS"], g["qE"]]; boff += 2 * g["tLen"]
        expect.append(g["blocks"])
    gb["band"] = np.concatenate(band + [np.zeros(4, np.int32)]).astype(np.int32)
    return gb, expect


def pack_segments(recs):
    """Whole-function batch (lra_b200_indel_refine_batch) over captured segments that share k / scoring / endAlign."""
    r0 = recs[0]
    assert all((r["k"], r["match"], r["mismatch"], r["indel"], r["end_align"]) == (r0["k"], r0["match"], r0["mismatch"], r0["indel"], r0["end_align"]) for r in recs)
    q_parts, t_parts, qb, tb, qo, to = [], [], [], [], 0, 0
    for r in recs:
        qb.append(qo); q_parts.append(r["read"]); qo += len(r["read"])
        tb.append((to - r["t_win_off"]) & 0xFFFFFFFF); t_parts.append(r["twin"]); to += len(r["twin"])
    cnt = np.array([len(r["blocks_in"]) for r in recs], np.int32)
    off = np.zeros(len(recs), np.uint64); off[1:] = np.cumsum(cnt[:-1])
    return dict(q_arena=np.frombuffer(b"".join(q_parts) + b"N" * 16, dtype=np.uint8).copy(),
                t_arena=np.frombuffer(b"".join(t_parts) + b"N" * 16, dtype=np.uint8).copy(),
                blocks_in=np.concatenate([r["blocks_in"] for r in recs]).astype(np.uint32), blk_off=off, blk_cnt=cnt,
                q_base=np.array(qb, np.uint32), t_base=np.array(tb, np.uint32),
                read_len=np.array([len(r["read"]) for r in recs], np.int32), contig_len=np.array([r["contig_len"] for r in recs], np.int32),
                k=r0["k"], match=r0["match"], mismatch=r0["mismatch"], indel=r0["indel"], end_align=r0["end_align"])
