"""Golden loss values, generated in the authoring container from the REFERENCE's own code:

  * Node2Vec.loss(model, edges) of /root/reference/ADSCModel/node_embeddings.py:26-31, imported in place, evaluated on the
    karate tables recorded in golden_karate.npz (the reference's own run): at initialisation, after the pre-training o1
    epoch and after the first full o1+o2 iteration;
  * the o2 positive-pair objective has NO reference counterpart: the fixture records an independent
    float64 numpy evaluation of its definition  sum_{walk, i, j in window(i), j != i} -log sigmoid(node[w_j] . ctx[w_i])
    on the same tables and walks;
  * Community2Vec.loss of the reference raises (`model.vocab(x)`, community_embeddings.py:48): recorded as such.

    python tests/golden/make_golden_losses.py        -> tests/golden/golden_losses.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    O.load_ref("tuned", with_python_sources=True)
    from ADSCModel.node_embeddings import Node2Vec
    from utils.embedding import Vocab
    g = np.load(os.path.join(HERE, "golden_karate.npz"))
    edges, walks = g["edges"], g["walks_ids"]

    class M(object):
        pass
    model = M()
    model.vocab = {}
    for i, c in enumerate(g["degrees"]):
        v = Vocab()
        v.count, v.index, v.sample_probability = int(c), i, 1.0
        model.vocab[i + 1] = v
    out = {"source": "reference Node2Vec.loss (node_embeddings.py:26-31) imported in place; o2 objective: float64 numpy",
           "node2vec_loss": {}, "o2_pos_loss": {}}
    learner = Node2Vec(workers=1, negative=4, lr=0.1)
    for stage in ("init_node", "pre_o1_node", "it0_o2_node"):
        model.node_embedding = g[stage]
        out["node2vec_loss"][stage] = float(learner.loss(model, edges))
    try:
        from ADSCModel.community_embeddings import Community2Vec
        model.k = 2
        Community2Vec(model, lr=0.1).loss(list(range(1, 35)), model, 0.01)
        out["community2vec_loss"] = "ran"
    except Exception as e:  # the reference's own code path is broken
        out["community2vec_loss"] = "reference raises %s" % type(e).__name__
    W = 3
    for stage in ("pre", "it0"):
        node, ctx = g[stage + "_o2_node"].astype(np.float64), g[stage + "_o2_ctx"].astype(np.float64)
        tot, cnt = 0.0, 0
        for w in walks:
            rows = np.asarray(w, np.int64) - 1
            for i in range(len(rows)):
                for j in range(max(0, i - W), min(len(rows), i + W + 1)):
                    if j != i:
                        z = float(node[rows[j]] @ ctx[rows[i]])
                        tot += np.log1p(np.exp(-z)) if z > 0 else (-z + np.log1p(np.exp(z)))
                        cnt += 1
        out["o2_pos_loss"][stage] = {"window": W, "sum": tot, "pairs": cnt}
    with open(os.path.join(HERE, "golden_losses.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
