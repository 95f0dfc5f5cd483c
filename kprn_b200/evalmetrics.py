"""Output side of the path: the score file the scorer writes and the reference's ranking metrics,
with its exact semantics, so "hit@10 within +-0.001" is checked the way the reference computes it.
Paths relative to /root/reference/release/songPathRnn/.

  write_scores   eval/test_from_checkpoint.lua:109-116   "counter \\t %.5f score \\t label"
  combine        eval/combine_result.py:27               "user \\t item \\t label \\t score"
  resort         eval/resort.py:38                       sort by (user, -score), users restricted to test_user_id.txt
  eval_one_rating / evaluate   eval/eval_score.py:34-46,123-160   hit@k / ndcg@k, k = 1..15
"""
import heapq
import math


def format_score(x):
    return "%.5f" % float(x)          # string.format("%.5f", score), test_from_checkpoint.lua:112


def write_scores(path, scores, labels, start=0):
    with open(path, "w") as f:
        for i, (s, l) in enumerate(zip(scores, labels)):
            lab = ("%d" % l) if float(l).is_integer() else repr(float(l))
            f.write("%d\t%s\t%s\n" % (start + i, format_score(s), lab))


def combine(entity_lines, result_lines):
    """entity line: idx \\t user \\t item ... ; result line: idx \\t score \\t label -> (user, item, label, score-string)."""
    out = []
    for e, r in zip(entity_lines, result_lines):
        el, rl = e.strip().split("\t"), r.strip().split("\t")
        out.append((el[1], el[2], rl[-1], rl[-2]))
    return out


def resort(rows, user_ids):
    """rows of (user, item, label, score); keeps users in user_ids; stable sort by (int(user), -score)."""
    keep = set(user_ids)
    items = [(int(u), it, lab, float(sc)) for (u, it, lab, sc) in rows if u in keep]
    return sorted(items, key=lambda x: (x[0], -x[-1]))


def eval_one_rating(i_gnd, i_pre, K):
    """eval_score.py:34-46 verbatim semantics: dict keyed by position (positive first), heapq.nlargest -> ties favour
    the lower index, i.e. the positive; all-zero predictions score (0, 0)."""
    if sum(i_pre) == 0:
        return 0, 0
    map_score = {item: score for item, score in enumerate(i_pre)}
    target_item = i_gnd.index(1.0)
    rank_list = heapq.nlargest(K, map_score, key=map_score.get)
    hit = 1 if target_item in rank_list else 0
    ndcg = 0
    for i, item in enumerate(rank_list):
        if item == target_item:
            ndcg = math.log(2) / math.log(i + 2)
            break
    return hit, ndcg


def evaluate(sorted_rows, test_samples, max_k=15):
    """sorted_rows: output of resort(); test_samples: iterable of (user, pos_item, [neg_items]).
    Returns (hit[k], ndcg[k]) averaged over the samples whose negatives were all scored (eval_score.py:95-160)."""
    by_pair = {}
    for (u, it, lab, sc) in sorted_rows:
        by_pair[(str(u), str(it))] = float(sc)
    hits, ndcgs = [], []
    for user, pos, negs in test_samples:
        user = str(user)
        if (user, str(pos)) not in by_pair or any((user, str(n)) not in by_pair for n in negs):
            continue
        gnd = [1.0] + [0.0] * len(negs)
        pre = [by_pair[(user, str(pos))]] + [by_pair[(user, str(n))] for n in negs]
        h, d = [], []
        for k in range(1, max_k + 1):
            hk, dk = eval_one_rating(gnd, pre, k)
            h.append(hk)
            d.append(dk)
        hits.append(h)
        ndcgs.append(d)
    n = max(len(hits), 1)
    hit_avg = [sum(x[i] for x in hits) / n for i in range(max_k)]
    ndcg_avg = [sum(x[i] for x in ndcgs) / n for i in range(max_k)]
    return hit_avg, ndcg_avg, len(hits)


def format_result(hit_avg, ndcg_avg):
    """eval_res_<alpha>.txt lines (eval_score.py:156-160)."""
    return ("hit scores\t:" + "\t".join("%.5f" % v for v in hit_avg) + "\n" +
            "ndcg scores\t:" + "\t".join("%.5f" % v for v in ndcg_avg) + "\n")
