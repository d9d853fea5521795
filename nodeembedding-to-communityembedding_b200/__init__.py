"""comemb_b200 -- B200-native SGD hot path of ComEmb (o1 / o2 / o3 + walks) behind the reference's entry points.

Layout mirrors the reference so call sites change only their import root:
    utils.training_sdg_inner  -> comemb_b200.utils.training_sdg_inner   (train_o1, train_o2, train_sg, init, FAST_VERSION)
    ADSCModel.node_embeddings -> comemb_b200.ADSCModel.node_embeddings   (Node2Vec) ... etc.
The arithmetic lives in csrc/ (CUDA, sm_100a) behind the C ABI of include/comemb_b200.h.
"""
__version__ = "0.1.0"
