"""``import pointops`` drop-in: re-exports pointcloudpdf_b200.pointops under the reference's
package name (libs/pointops/__init__.py:1), so the unmodified
pointcept/models/point_transformer and pointcept/recognizers pick up the B200 kernels."""
from pointcloudpdf_b200.pointops import *  # noqa: F401,F403
from pointcloudpdf_b200.pointops import (  # noqa: F401
    KNNQuery, FarthestPointSampling, Grouping, Interpolation, Subtraction, Aggregation,
    furthestsampling, knnquery, queryandgroup, clear_caches, set_cache_sizes, register_host_offset,
)
