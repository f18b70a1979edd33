"""Inputs for the a8 test (SplitRoughClustersWithGaps): anchor lists whose rough clusters are runs of anchors along a diagonal with gaps below and above
RoughClustermaxGap, short pieces (dropped), pieces that come back close to the previous one (re-joined), repetitive clusters (anchorfreq >= 10: kept whole),
both strands, with and without contig indices on the rough clusters."""
import numpy as np


def rough_list(rng, maxGap=1000, K=17, with_chrom=False):
    n_rough = int(rng.integers(1, 7))
    Q, T, rs, re_, box, st, fr, ch = [], [], [], [], [], [], [], []
    for c in range(n_rough):
        strand = int(rng.random() < 0.3)
        q = int(rng.integers(0, 5000)); t = int(rng.integers(100_000, 2_000_000))
        pts = []
        for _ in range(int(rng.integers(1, 6))):                 # pieces
            for _ in range(int(rng.choice([1, 2, 3, 6, 15]))):
                step = int(rng.integers(1, 120))
                q += step; t += (step + int(rng.integers(-5, 6))) * (1 if strand == 0 else -1)
                pts.append((q, max(t, 0)))
            g = int(rng.choice([200, maxGap - 1, maxGap, maxGap + 1, maxGap + 5, maxGap + 12, 3 * maxGap, 20 * maxGap]))
            jit = 0 if g in (maxGap + 5, maxGap + 12) else int(rng.integers(-300, 301))     # just over the gap on the diagonal: split, then re-joined
            q += g; t += (g + jit) * (1 if strand == 0 else -1)
        a = np.array(pts, np.int64)
        a = a[np.lexsort((a[:, 1], a[:, 0]))]                    # CartesianSort: q, then t
        rs.append(len(Q)); Q += a[:, 0].tolist(); T += a[:, 1].tolist(); re_.append(len(Q))
        box.append([int(a[:, 0].min()), int(a[:, 0].max()) + K, int(a[:, 1].min()), int(a[:, 1].max()) + K])
        st.append(strand); fr.append(float(rng.choice([1.0, 1.5, 9.9, 10.0, 12.5]))); ch.append(int(rng.integers(0, 3)) if with_chrom else -1)
    if rng.random() < 0.2:                                       # an empty rough cluster
        rs.append(len(Q)); re_.append(len(Q)); box.append([0, 0, 0, 0]); st.append(0); fr.append(1.0); ch.append(-1)
    return np.array(Q, np.uint32), np.array(T, np.uint32), dict(start=np.array(rs, np.int32), end=np.array(re_, np.int32), box=np.array(box, np.uint32),
                                                                 strand=np.array(st, np.uint8), freq=np.array(fr, np.float32), chrom=np.array(ch, np.int32))
