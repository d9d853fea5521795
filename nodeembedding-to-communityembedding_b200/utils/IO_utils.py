"""Label / embedding text formats of the reference (/root/reference/utils/IO_utils.py), kept for interop."""
from os import makedirs
from os.path import dirname, join as path_join

import numpy as np


def load_ground_true(path="data/", file_name=None, multilabel=False):
    """`<node_id>\\t<label>` lines -> (labels ordered by node id, number of communities = max label)
    (IO_utils.py:19-47)."""
    labels, kmax = {}, 0
    with open(path_join(path, file_name + ".labels"), "r") as f:
        for line in f:
            tok = line.strip().split("\t")
            if len(tok) < 2:
                continue
            node, lab = int(tok[0]), int(tok[1])
            kmax = max(kmax, lab)
            labels.setdefault(node, []).append(lab)
    ret = [labels[k] if multilabel else labels[k][0] for k in sorted(labels)]
    return ret, kmax


def save_embedding(embeddings, file_name, path="data"):
    """`<node_id>\\t<v1> <v2> ...` with 1-based ids in row order (IO_utils.py:49-62)."""
    if hasattr(embeddings, "detach"):
        embeddings = embeddings.detach().cpu().numpy()
    full = path_join(path, file_name + ".txt")
    makedirs(dirname(full) or ".", exist_ok=True)
    with open(full, "w") as f:
        for i, row in enumerate(embeddings):
            f.write(str(i + 1) + "\t" + " ".join(str(v) for v in row) + "\n")


def load_embedding(file_name, path="data", ext=".txt"):
    ret = []
    with open(path_join(path, file_name + ext), "r") as f:
        for line in f:
            tok = line.strip().split("\t")
            ret.append([float(v) for v in tok[1].strip().split(" ")])
    return np.array(ret, dtype=np.float32)
