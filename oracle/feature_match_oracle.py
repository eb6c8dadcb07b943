"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's feature-space correspondence search,
data_preprocess/3DMatch_Feature.py:158-166 -- only tests/ may import this.

    distance = np.sqrt(2 - 2 * (src_desc @ tgt_desc.T) + 1e-6)
    source_idx = np.argmin(distance, axis=1); source_dis = np.min(distance, axis=1)
    if use_mutual: target_idx = np.argmin(distance, axis=0); mutual_nearest = target_idx[source_idx] == arange
"""
import numpy as np


def distance_matrix(src_desc, tgt_desc):
    src_desc = np.asarray(src_desc, dtype=np.float32); tgt_desc = np.asarray(tgt_desc, dtype=np.float32)
    return np.sqrt(2 - 2 * (src_desc @ tgt_desc.T) + 1e-6)          # 3DMatch_Feature.py:158 (float32 throughout)


def correspondences(src_desc, tgt_desc, use_mutual=False):
    distance = distance_matrix(src_desc, tgt_desc)
    source_idx = np.argmin(distance, axis=1)                         # :159
    source_dis = np.min(distance, axis=1)                            # :160
    if use_mutual:                                                   # :161-164
        target_idx = np.argmin(distance, axis=0)
        mutual_nearest = (target_idx[source_idx] == np.arange(source_idx.shape[0]))
        corr = np.concatenate([np.where(mutual_nearest == 1)[0][:, None], source_idx[mutual_nearest][:, None]], axis=-1)
    else:                                                            # :165-166
        corr = np.concatenate([np.arange(source_idx.shape[0])[:, None], source_idx[:, None]], axis=-1)
    return corr, source_idx, source_dis, distance
