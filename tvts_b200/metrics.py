"""Retrieval metrics of the validation loop, restated in numpy from the reference:
  t2v_metrics / v2t_metrics / cols2metrics   v2/model/metric.py:16-125, 125-215, 285-295
(host-side numpy like the reference's; the embeddings they consume come from the CUDA towers).  Rank of a ground-truth item =
number of candidates with a strictly smaller distance; ties are broken "optimistically" for text->video (reference :64) and by
averaging the tied positions for video->text (reference :156)."""
import numpy as np


def cols2metrics(cols, num_queries):
    cols = np.asarray(cols, dtype=np.float64)
    m = {
        "R1": 100 * float(np.sum(cols == 0)) / num_queries,
        "R5": 100 * float(np.sum(cols < 5)) / num_queries,
        "R10": 100 * float(np.sum(cols < 10)) / num_queries,
        "R50": 100 * float(np.sum(cols < 50)) / num_queries,
        "MedR": float(np.median(cols) + 1),
        "MeanR": float(np.mean(cols) + 1),
    }
    stats = np.array([m["R1"], m["R5"], m["R10"]], dtype=np.float64)
    m["geometric_mean_R1-R5-R10"] = float(np.exp(np.log(stats).mean())) if np.all(stats > 0) else 0.0
    return m


def t2v_metrics(sims, query_masks=None):
    """sims [num_queries, num_vids]: sims[i, j] = <text_i, video_j>; query i belongs to video i // (num_queries // num_vids)."""
    sims = np.asarray(sims)
    assert sims.ndim == 2, "expected a matrix"
    nq, nv = sims.shape
    dists = -sims
    qpv = nq // nv
    gt = dists[np.arange(nq), np.arange(nq) // qpv]
    cols = (np.sort(dists, axis=1) < gt[:, None]).sum(1)          # first position of the ground-truth distance (optimistic ties)
    if query_masks is not None:
        keep = np.asarray(query_masks).reshape(-1).astype(bool)
        cols = cols[keep]
        nq = int(keep.sum())
    return cols2metrics(cols, nq)


def v2t_metrics(sims, query_masks=None):
    """sims as for t2v_metrics; the rank of a video is the best (smallest) rank among its captions, ties averaged."""
    sims = np.asarray(sims).T
    nq, nc = sims.shape
    dists = -sims.astype(np.float64)
    cpv = nc // nq
    missing = 1e8
    ranks = []
    for i in range(nq):
        row = dists[i].copy()
        if query_masks is not None:
            row[np.logical_not(np.asarray(query_masks).reshape(-1))] = missing
        srt = np.sort(row)
        best = np.inf
        for j in range(i * cpv, (i + 1) * cpv):
            if row[j] == missing:
                continue
            pos = np.where(srt == row[j])[0]
            best = min(best, pos.mean())
        ranks.append(best)
    return cols2metrics(np.array(ranks), nq)
