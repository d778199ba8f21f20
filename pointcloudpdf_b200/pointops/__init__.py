"""Drop-in for the reference package ``pointops`` (libs/pointops/functions/__init__.py:1-14).

Same names, positional signatures, defaults and return dtypes; every operator runs a
hand-written sm_100a kernel from libpointops_b200.so (no Triton, no CPU fallback).  Legacy
pointops2 spellings used by BASELINE.json's north_star (furthestsampling, knnquery,
queryandgroup; libs/pointops2/functions/pointops.py:34,56,964) are provided as aliases.
"""
from .query import knn_query, ball_query, random_ball_query, KNNQuery, BallQuery, RandomBallQuery
from .sampling import farthest_point_sampling, FarthestPointSampling
from .grouping import grouping, grouping2, Grouping, grouping_split
from .interpolation import interpolation, interpolation2, Interpolation
from .subtraction import subtraction, Subtraction
from .aggregation import aggregation, Aggregation
from .attention import attention_relation_step, attention_fusion_step, AttentionRelationStep, AttentionFusionStep
from .utils import (
    query_and_group,
    knn_query_and_group,
    ball_query_and_group,
    batch2offset,
    offset2batch,
)
from ._common import clear_caches, set_cache_sizes, register_host_offset

# legacy pointops2 names (note knnquery's different argument order)
furthestsampling = farthest_point_sampling


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """pointops2.knnquery(nsample, xyz, new_xyz, offset, new_offset)
    (libs/pointops2/functions/pointops.py:37-56)."""
    if new_xyz is None:
        new_xyz = xyz
    return knn_query(nsample, xyz, offset, new_xyz, new_offset)


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """pointops2.queryandgroup (libs/pointops2/functions/pointops.py:964-1001): kNN + gather,
    returns only the grouped tensor."""
    out, _ = query_and_group(nsample, xyz, new_xyz, feat, idx, offset, new_offset, with_xyz=use_xyz)
    return out
